/* empty shim for Code/Convex_Tree.h:13 */
