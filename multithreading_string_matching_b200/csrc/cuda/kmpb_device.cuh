// kmpb_device.cuh -- context layout and launch entry points shared by the .cu files of libkmpb200.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../host/kmpb_internal.h"

#define KMPB_CUDA(call)                                                                          \
    do {                                                                                         \
        cudaError_t e_ = (call);                                                                 \
        if (e_ != cudaSuccess)                                                                   \
            return kmpb_fail(KMPB_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                             __FILE__, __LINE__);                                                \
    } while (0)

constexpr int KMPB_COPY_STREAMS = 4;          // H2D chunk pipeline depth of kmpb_count_host
constexpr uint32_t KMPB_SMEM_COUNTS_MAX = 4096; // distinct patterns counted in shared memory; above: global atomics

// Device-resident tables of one pattern set (device pointers).
struct kmpb_device_tables {
    uint32_t n_pat = 0, n_uniq = 0, n_class = 0, n_state = 0, max_len = 0, min_len = 0;
    // distinct patterns
    uint8_t *uniq_blob = nullptr;   // [sum len]
    uint32_t *uniq_off = nullptr;   // [n_uniq+1]
    uint32_t *uniq_len = nullptr;   // [n_uniq]
    uint32_t *pat_to_uniq = nullptr; // [n_pat]
    // per-pattern engine: failure tables and byte-indexed KMP DFAs, both built on the device
    int32_t *pi = nullptr;          // [sum len], pattern u at uniq_off[u]
    uint8_t *perpat_dfa = nullptr;  // pattern u: uniq_len[u] rows of 256 entries at 256*uniq_off[u]
    // union engine
    uint32_t *vtab = nullptr;       // [vtab_words] start-anchored verification tables (automaton.c)
    uint32_t *filter = nullptr;     // [256]
};

struct kmpb_ctx {
    int device = 0;
    int sm_count = 0;
    int engine = KMPB_ENGINE_AUTO;
    size_t smem_optin = 0;
    cudaStream_t stream = nullptr;                      // default stream of the context
    cudaStream_t copy_stream[KMPB_COPY_STREAMS] = {};   // H2D + kernel pipeline of kmpb_count_host
    cudaEvent_t ev[4] = {};
    cudaEvent_t ev_kernel[2] = {};                      // around the dominant kernel when profile is on
    bool profile = false;
    kmpb_tables host = {};
    bool have_tables = false;
    kmpb_device_tables dev;
    // scratch, grown on demand
    uint64_t *d_uniq_counts = nullptr;   // [n_uniq] per-launch accumulators (per stream slot)
    uint32_t *d_work = nullptr;          // work counters (per stream slot)
    uint32_t *d_items = nullptr;         // item -> first packet table (per stream slot)
    size_t items_cap = 0;                // entries per slot
    uint8_t *d_stage_bytes[KMPB_COPY_STREAMS] = {};
    uint64_t *d_stage_off[KMPB_COPY_STREAMS] = {};
    size_t stage_bytes_cap = 0, stage_off_cap = 0;
    // pinned twins of the staging slots for the streamed savefile path (kmpb_count_pcap)
    uint8_t *h_stage_bytes[KMPB_COPY_STREAMS] = {};
    uint64_t *h_stage_off[KMPB_COPY_STREAMS] = {};
    size_t h_stage_bytes_cap = 0, h_stage_off_cap = 0;
    cudaEvent_t ev_h2d[KMPB_COPY_STREAMS] = {};
    uint64_t *d_counts = nullptr;        // [n_pat] result of kmpb_count_host
    bool attr_union_set = false, attr_perpat_set = false; // per-device function attributes
    uint64_t launches = 0;
    double last_ms[2] = {0, 0};
};

// One batch on the device.  d_bytes[0] is absolute byte `abs_base` of the CSR byte space (a multiple
// of 512); offsets are absolute.  All launches go to `stream`; slot selects the scratch set.
struct kmpb_batch {
    const uint8_t *d_bytes;
    uint64_t abs_base;
    const uint64_t *d_offsets; // [n_packets+1]
    uint64_t n_packets;
    uint64_t first_byte, end_byte; // offsets[0], offsets[n_packets] (host copies)
};

// Where the union kernel's last block adds the per-pattern counts (file order) itself: up to 8 count
// vectors, this GPU's and/or peers' mapped over NVLink.  n = 0: the caller expands the counts afterwards.
constexpr int KMPB_MAX_OUT = 8;
struct kmpb_fused_out {
    unsigned long long *vec[KMPB_MAX_OUT] = {};
    uint32_t n = 0;
};

// tables.cu
int kmpb_upload_tables(kmpb_ctx *ctx);
void kmpb_release_tables(kmpb_ctx *ctx);
// perpat_kernel.cu
int kmpb_launch_perpat(kmpb_ctx *ctx, const kmpb_batch &b, uint64_t *d_uniq_counts, cudaStream_t stream);
// union_kernel.cu
int kmpb_launch_union(kmpb_ctx *ctx, const kmpb_batch &b, int slot, uint64_t *d_uniq_counts, cudaStream_t stream,
                      const kmpb_fused_out &out = kmpb_fused_out());
int kmpb_union_scratch(kmpb_ctx *ctx, uint64_t max_batch_bytes);
// api.cu
int kmpb_launch_expand(kmpb_ctx *ctx, const uint64_t *d_uniq_counts, uint64_t *d_counts, cudaStream_t stream);
