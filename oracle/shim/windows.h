/* Test-infrastructure shim (NOT product code): lets the reference's Windows-only timer
 * header (Code/cyw_timer.h:29-35) and itoa use (Code/basic_functions.h:316) compile on Linux
 * when building oracle/_ref from the sources under /root/reference. */
#ifndef SCHT_SHIM_WINDOWS_H
#define SCHT_SHIM_WINDOWS_H
#include <time.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <math.h>
typedef union { long long QuadPart; } LARGE_INTEGER;
static inline int QueryPerformanceFrequency(LARGE_INTEGER* f) { f->QuadPart = 1000000000LL; return 1; }
static inline int QueryPerformanceCounter(LARGE_INTEGER* c) {
    struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
    c->QuadPart = (long long)ts.tv_sec * 1000000000LL + ts.tv_nsec; return 1;
}
static inline char* itoa(int v, char* buf, int base) { (void)base; sprintf(buf, "%d", v); return buf; }
#endif
