// gather_bench.cu -- random-32-byte-sector gather microbenchmark.
//
// The classify kernel is bound by how many independent random 32-B sectors HBM
// can deliver per second (one sector per set probe), not by streaming bandwidth.
// This kernel measures that ceiling on the same box: every thread keeps
// `loads_in_flight` independent loads outstanding at pseudo-random
// sector-aligned addresses of a buffer much larger than L2.
//   variant 0: one 256-bit load  (ld.global.nc.L1::no_allocate.v4.u64) -- what the probe uses
//   variant 1: two 128-bit loads of the same sector
//   variant 2: one 64-bit load per sector (touches the sector, moves 8 bytes to the SM)
#include "launchers.h"

namespace {

__device__ __forceinline__ uint32_t mix32(uint32_t h)
{
    h ^= h >> 16;
    h *= 0x7FEB352Du;
    h ^= h >> 15;
    h *= 0x846CA68Bu;
    h ^= h >> 16;
    return h;
}

template <int MLP, int VARIANT>
__global__ void k_gather(const tb_bucket *__restrict__ buf, uint64_t n_sectors, int steps,
                         unsigned long long *sink)
{
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long acc = 0;
    uint32_t state = mix32(tid * 2654435761u + 12345u);
    for (int s = 0; s < steps; s++) {
        uint64_t idx[MLP];
#pragma unroll
        for (int u = 0; u < MLP; u++) {
            state = state * 1664525u + 1013904223u;
            uint32_t h1 = mix32(state);
            uint32_t h2 = mix32(state ^ 0x9E3779B9u);
            // 64-bit uniform index so buffers > 2^32 sectors (128 GB) still work
            uint64_t h = ((uint64_t)h1 << 32) | h2;
            idx[u] = __umul64hi(h, n_sectors);
        }
        if (VARIANT == 0) {
            tb_bucket v[MLP];
#pragma unroll
            for (int u = 0; u < MLP; u++) v[u] = tb_load_bucket(buf + idx[u]);
#pragma unroll
            for (int u = 0; u < MLP; u++) acc ^= v[u].s[0] ^ v[u].s[1] ^ v[u].s[2] ^ v[u].s[3];
        } else if (VARIANT == 1) {
            ulonglong2 v[MLP][2];
#pragma unroll
            for (int u = 0; u < MLP; u++) {
                const ulonglong2 *p = reinterpret_cast<const ulonglong2 *>(buf + idx[u]);
                asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0,%1}, [%2];"
                             : "=l"(v[u][0].x), "=l"(v[u][0].y) : "l"(p));
                asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0,%1}, [%2];"
                             : "=l"(v[u][1].x), "=l"(v[u][1].y) : "l"(p + 1));
            }
#pragma unroll
            for (int u = 0; u < MLP; u++) acc ^= v[u][0].x ^ v[u][0].y ^ v[u][1].x ^ v[u][1].y;
        } else {
            unsigned long long v[MLP];
#pragma unroll
            for (int u = 0; u < MLP; u++) {
                asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];"
                             : "=l"(v[u]) : "l"(buf + idx[u]));
            }
#pragma unroll
            for (int u = 0; u < MLP; u++) acc ^= v[u];
        }
    }
    if (acc == 0x1234567812345678ull) sink[0] = acc;  // never true in practice; defeats DCE
}

template <int VARIANT>
cudaError_t launch_v(const tb_bucket *buf, uint64_t n, int mlp, int tpb, int blocks, int steps,
                     unsigned long long *sink, cudaStream_t st)
{
    switch (mlp) {
    case 1: k_gather<1, VARIANT><<<blocks, tpb, 0, st>>>(buf, n, steps, sink); break;
    case 2: k_gather<2, VARIANT><<<blocks, tpb, 0, st>>>(buf, n, steps, sink); break;
    case 4: k_gather<4, VARIANT><<<blocks, tpb, 0, st>>>(buf, n, steps, sink); break;
    case 8: k_gather<8, VARIANT><<<blocks, tpb, 0, st>>>(buf, n, steps, sink); break;
    case 16: k_gather<16, VARIANT><<<blocks, tpb, 0, st>>>(buf, n, steps, sink); break;
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace

cudaError_t tb_launch_gather(const void *d_buf, uint64_t n_sectors, int loads_in_flight,
                             int threads_per_block, int blocks, int steps, int variant,
                             unsigned long long *d_sink, cudaStream_t stream)
{
    const tb_bucket *buf = reinterpret_cast<const tb_bucket *>(d_buf);
    switch (variant) {
    case 0: return launch_v<0>(buf, n_sectors, loads_in_flight, threads_per_block, blocks, steps, d_sink, stream);
    case 1: return launch_v<1>(buf, n_sectors, loads_in_flight, threads_per_block, blocks, steps, d_sink, stream);
    case 2: return launch_v<2>(buf, n_sectors, loads_in_flight, threads_per_block, blocks, steps, d_sink, stream);
    default: return cudaErrorInvalidValue;
    }
}
