// synth.cu -- counter-based synthetic UDP payload streams (BASELINE.json configs 3-5).
//
// Packet p of a stream is a pure function of (seed, p): the same code runs on the host (for the CPU
// baselines, which need a pcap file) and on the device (for device-resident benchmarks: no 14 GB
// H2D copy), and any slice [first, first+count) of the stream can be produced on its own.
//   - payload bytes: printable ASCII 0x20..0x7e from a splitmix64 counter hash, 8 bytes per hash;
//   - the last byte of every payload is 0x00, so the reference's strlen() scan (serial.c:191) is
//     well defined when serial.c itself is the comparator;
//   - `plants` patterns from the plant set overwrite pseudo-random positions (later plants win).
#include <algorithm>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "kmpb_device.cuh"

namespace {

__host__ __device__ inline uint64_t mix64(uint64_t x)
{
    x += 0x9e3779b97f4a7c15ull;
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
    return x ^ (x >> 31);
}

__host__ __device__ inline uint32_t synth_len(uint64_t seed, uint32_t payload_len, uint32_t len_mode, uint64_t p)
{
    if (len_mode == 0) return payload_len;
    const uint32_t r = (uint32_t)(mix64(seed ^ mix64(p ^ 0x4c454e4754485f5full)) % 100u);
    return r < 40 ? 64u : r < 60 ? 576u : r < 90 ? 1400u : 9000u; // BASELINE config 5 mix
}

struct plant_set {
    const uint8_t *blob;
    const uint32_t *off;
    uint32_t n, plants;
};

// 8 payload bytes: word j of packet p whose payload has L bytes
__host__ __device__ inline void synth_word(uint64_t seed, uint64_t p, uint32_t j, uint32_t L, const plant_set &ps,
                                           uint8_t out[8])
{
    const uint64_t h = mix64(seed ^ mix64(p * 0x632be59bd9b4e019ull + j));
    const uint32_t i0 = j * 8u;
    for (uint32_t k = 0; k < 8; k++) out[k] = (uint8_t)(0x20u + ((((uint32_t)(h >> (8 * k)) & 0xffu) * 95u) >> 8));
    for (uint32_t t = 0; t < ps.plants && ps.n; t++) {
        const uint64_t hp = mix64(seed ^ mix64(p * 2u + 1u) ^ (0x504c414e54ull + t));
        const uint32_t which = (uint32_t)(hp % ps.n);
        const uint32_t m = ps.off[which + 1] - ps.off[which];
        if (L < m + 1u) continue; // keep the final NUL intact
        const uint32_t at = (uint32_t)((hp >> 32) % (uint64_t)(L - m));
        for (uint32_t k = 0; k < 8; k++) {
            const uint32_t i = i0 + k;
            if (i >= at && i < at + m) out[k] = ps.blob[ps.off[which] + (i - at)];
        }
    }
    if (L && i0 + 8u >= L && i0 < L) out[L - 1u - i0] = 0;
}

__global__ void kmpb_synth_kernel(uint64_t seed, uint32_t payload_len, uint32_t len_mode, plant_set ps,
                                  uint64_t first, uint64_t count, const uint64_t *__restrict__ offsets,
                                  uint8_t *__restrict__ bytes)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t q = warp; q < count; q += n_warps) {
        const uint64_t p = first + q;
        const uint32_t L = synth_len(seed, payload_len, len_mode, p);
        uint8_t *dst = bytes + offsets[q];
        const uint32_t words = (L + 7u) / 8u;
        for (uint32_t j = lane; j < words; j += 32) {
            uint8_t w[8];
            synth_word(seed, p, j, L, ps, w);
            uint8_t *d = dst + 8u * j;
            if ((((uintptr_t)d) & 7) == 0 && 8u * j + 8u <= L) {
                uint64_t v;
                memcpy(&v, w, 8);
                *reinterpret_cast<uint64_t *>(d) = v;
            } else {
                for (uint32_t k = 0; k < 8 && 8u * j + k < L; k++) d[k] = w[k];
            }
        }
    }
}

int check_cfg(const kmpb_synth *cfg)
{
    if (cfg == nullptr) return kmpb_fail(KMPB_EINVAL, "NULL synth config");
    if (cfg->plants && cfg->n_plant && (cfg->plant_blob == nullptr || cfg->plant_off == nullptr))
        return kmpb_fail(KMPB_EINVAL, "plants requested without a plant set");
    return KMPB_OK;
}

} // namespace

extern "C" {

uint64_t kmpb_synth_bytes(const kmpb_synth *cfg, uint64_t first, uint64_t count)
{
    if (cfg == nullptr) return 0;
    if (cfg->len_mode == 0) return count * (uint64_t)cfg->payload_len;
    uint64_t total = 0;
    for (uint64_t q = 0; q < count; q++) total += synth_len(cfg->seed, cfg->payload_len, cfg->len_mode, first + q);
    return total;
}

// offsets relative to the slice start; usable without a GPU
static void synth_offsets(const kmpb_synth *cfg, uint64_t first, uint64_t count, uint64_t *offsets)
{
    uint64_t at = 0;
    for (uint64_t q = 0; q < count; q++) {
        offsets[q] = at;
        at += synth_len(cfg->seed, cfg->payload_len, cfg->len_mode, first + q);
    }
    offsets[count] = at;
}

int kmpb_synth_fill_host(const kmpb_synth *cfg, uint64_t first, uint64_t count, uint8_t *bytes, uint64_t *offsets)
{
    int rc = check_cfg(cfg);
    if (rc) return rc;
    if ((count && bytes == nullptr) || offsets == nullptr) return kmpb_fail(KMPB_EINVAL, "NULL output buffer");
    synth_offsets(cfg, first, count, offsets);
    plant_set ps{cfg->plant_blob, cfg->plant_off, cfg->plants ? cfg->n_plant : 0, cfg->plants};
#pragma omp parallel for schedule(static)
    for (int64_t q = 0; q < (int64_t)count; q++) {
        const uint32_t L = (uint32_t)(offsets[q + 1] - offsets[q]);
        uint8_t *dst = bytes + offsets[q];
        for (uint32_t j = 0; 8u * j < L; j++) {
            uint8_t w[8];
            synth_word(cfg->seed, first + (uint64_t)q, j, L, ps, w);
            const uint32_t n = L - 8u * j < 8u ? L - 8u * j : 8u;
            memcpy(dst + 8u * j, w, n);
        }
    }
    return KMPB_OK;
}

int kmpb_synth_fill_device(kmpb_ctx *ctx, const kmpb_synth *cfg, uint64_t first, uint64_t count,
                           uint8_t *d_bytes, uint64_t *d_offsets)
{
    int rc = check_cfg(cfg);
    if (rc) return rc;
    if (ctx == nullptr || (count && d_bytes == nullptr) || d_offsets == nullptr)
        return kmpb_fail(KMPB_EINVAL, "NULL argument");
    KMPB_CUDA(cudaSetDevice(ctx->device));
    std::vector<uint64_t> offsets(count + 1);
    synth_offsets(cfg, first, count, offsets.data());
    KMPB_CUDA(cudaMemcpyAsync(d_offsets, offsets.data(), (count + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
    plant_set ps{nullptr, nullptr, 0, cfg->plants};
    uint8_t *d_blob = nullptr;
    uint32_t *d_off = nullptr;
    if (cfg->plants && cfg->n_plant) {
        const size_t blob_len = cfg->plant_off[cfg->n_plant];
        KMPB_CUDA(cudaMalloc((void **)&d_blob, blob_len ? blob_len : 1));
        KMPB_CUDA(cudaMalloc((void **)&d_off, (cfg->n_plant + 1) * sizeof(uint32_t)));
        KMPB_CUDA(cudaMemcpyAsync(d_blob, cfg->plant_blob, blob_len, cudaMemcpyHostToDevice, ctx->stream));
        KMPB_CUDA(cudaMemcpyAsync(d_off, cfg->plant_off, (cfg->n_plant + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
        ps.blob = d_blob;
        ps.off = d_off;
        ps.n = cfg->n_plant;
    }
    if (count) {
        const int threads = 256;
        const uint64_t warps = count;
        const int grid = (int)std::min<uint64_t>((warps * 32 + threads - 1) / threads, (uint64_t)ctx->sm_count * 32);
        kmpb_synth_kernel<<<grid, threads, 0, ctx->stream>>>(cfg->seed, cfg->payload_len, cfg->len_mode, ps, first, count,
                                                            d_offsets, d_bytes);
        ctx->launches++;
        KMPB_CUDA(cudaGetLastError());
    }
    KMPB_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(d_blob);
    cudaFree(d_off);
    return KMPB_OK;
}

} // extern "C"
