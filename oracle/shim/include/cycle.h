/* Stand-in for the reference's missing include/cycle.h (qrdecomp.c:15 includes it,
 * the file is not in the reference tree).  Timing only; not part of the arithmetic. */
#ifndef ORACLE_SHIM_CYCLE_H
#define ORACLE_SHIM_CYCLE_H
#include <stdint.h>
#include <time.h>
typedef uint64_t ticks;
static inline ticks getticks(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (ticks)ts.tv_sec * 1000000000ull + (ticks)ts.tv_nsec;
}
#endif
