/* errors.c -- thread-local last-error message behind kmpb_last_error(). */
#include <stdarg.h>
#include <stdio.h>

#include "kmpb_internal.h"

static __thread char g_err[512];

int kmpb_fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

const char *kmpb_last_error(void) { return g_err; }

const char *kmpb_version(void) { return KMPB_VERSION; }

/* mpi_dumping.c:149-157: local_size[i] = N / P for every rank, rank 0 += N % P; displacements are
 * the running sum, so rank 0's slice comes first and is the long one. */
void kmpb_shard_range(uint64_t n_packets, uint32_t world, uint32_t rank, uint64_t *first, uint64_t *count)
{
    if (world == 0) world = 1;
    uint64_t base = n_packets / world, extra = n_packets % world;
    if (rank == 0) {
        *first = 0;
        *count = base + extra;
    } else {
        *first = extra + (uint64_t)rank * base;
        *count = rank < world ? base : 0;
    }
}
