/* empty shim for Code/basic_functions.h:14 */
