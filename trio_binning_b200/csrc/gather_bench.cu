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
//   variant 3/4: all 2/4 sectors of one random aligned 64-B / 128-B block (rate = blocks/s)
//   variant 13: one 32-bit load per sector, allocating in L1, evict_last in L2 -- the classify
//               kernel's bitmap load
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

// 32-byte load of one sector with assorted cache qualifiers (variants >= 5)
template <int VARIANT>
__device__ __forceinline__ tb_bucket load_q(const tb_bucket *p, uint64_t policy)
{
    tb_bucket b;
    if (VARIANT == 5)
        asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(b.s[0]), "=l"(b.s[1]), "=l"(b.s[2]), "=l"(b.s[3]) : "l"(p));
    else if (VARIANT == 6)
        asm volatile("ld.global.cv.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(b.s[0]), "=l"(b.s[1]), "=l"(b.s[2]), "=l"(b.s[3]) : "l"(p));
    else if (VARIANT == 7)
        asm volatile("ld.global.nc.L1::no_allocate.L2::64B.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(b.s[0]), "=l"(b.s[1]), "=l"(b.s[2]), "=l"(b.s[3]) : "l"(p));
    else if (VARIANT == 8)
        asm volatile("ld.global.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(b.s[0]), "=l"(b.s[1]), "=l"(b.s[2]), "=l"(b.s[3]) : "l"(p));
    else if (VARIANT == 9 || VARIANT == 10)
        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u64 {%0,%1,%2,%3}, [%4], %5;" : "=l"(b.s[0]), "=l"(b.s[1]), "=l"(b.s[2]), "=l"(b.s[3]) : "l"(p), "l"(policy));
    else if (VARIANT == 11)
        asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(b.s[0]), "=l"(b.s[1]), "=l"(b.s[2]), "=l"(b.s[3]) : "l"(p));
    else if (VARIANT == 12)
        asm volatile("ld.relaxed.gpu.global.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(b.s[0]), "=l"(b.s[1]), "=l"(b.s[2]), "=l"(b.s[3]) : "l"(p));
    else
        b = tb_load_bucket(p);
    return b;
}

template <int MLP, int VARIANT>
__global__ void k_gather(const tb_bucket *__restrict__ buf, uint64_t n_sectors, int steps,
                         unsigned long long *sink)
{
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long acc = 0;
    uint32_t state = mix32(tid * 2654435761u + 12345u);
    uint64_t policy = 0;
    if (VARIANT == 9)
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
    if (VARIANT == 10 || VARIANT == 13)
        asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(policy));
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
        if (VARIANT == 0 || VARIANT >= 5) {
            tb_bucket v[MLP];
#pragma unroll
            for (int u = 0; u < MLP; u++) v[u] = load_q<VARIANT>(buf + idx[u], policy);
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
        } else if (VARIANT == 13) {
            uint32_t v[MLP];
#pragma unroll
            for (int u = 0; u < MLP; u++)
                asm volatile("ld.global.nc.L2::cache_hint.u32 %0, [%1], %2;"
                             : "=r"(v[u]) : "l"(reinterpret_cast<const uint32_t *>(buf + idx[u]) + (state & 7u)), "l"(policy));
#pragma unroll
            for (int u = 0; u < MLP; u++) acc ^= v[u];
        } else if (VARIANT == 3 || VARIANT == 4) {
            // SPAN adjacent sectors of one aligned 64-B (3) or 128-B (4) block per unit:
            // does DRAM hand out the neighbours for free?
            constexpr int SPAN = VARIANT == 3 ? 2 : 4;
            tb_bucket v[MLP][SPAN];
#pragma unroll
            for (int u = 0; u < MLP; u++) {
                const tb_bucket *p = buf + (idx[u] & ~(uint64_t)(SPAN - 1));
#pragma unroll
                for (int q = 0; q < SPAN; q++) v[u][q] = tb_load_bucket(p + q);
            }
#pragma unroll
            for (int u = 0; u < MLP; u++)
#pragma unroll
                for (int q = 0; q < SPAN; q++) acc ^= v[u][q].s[0] ^ v[u][q].s[3];
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
    case 16:
        if (VARIANT == 3 || VARIANT == 4) return cudaErrorInvalidValue;  // register budget
        k_gather<((VARIANT == 3 || VARIANT == 4) ? 8 : 16), VARIANT><<<blocks, tpb, 0, st>>>(buf, n, steps, sink);
        break;
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
    case 3: return launch_v<3>(buf, n_sectors, loads_in_flight, threads_per_block, blocks, steps, d_sink, stream);
    case 4: return launch_v<4>(buf, n_sectors, loads_in_flight, threads_per_block, blocks, steps, d_sink, stream);
    case 5: return launch_v<5>(buf, n_sectors, loads_in_flight, threads_per_block, blocks, steps, d_sink, stream);
    case 6: return launch_v<6>(buf, n_sectors, loads_in_flight, threads_per_block, blocks, steps, d_sink, stream);
    case 7: return launch_v<7>(buf, n_sectors, loads_in_flight, threads_per_block, blocks, steps, d_sink, stream);
    case 8: return launch_v<8>(buf, n_sectors, loads_in_flight, threads_per_block, blocks, steps, d_sink, stream);
    case 9: return launch_v<9>(buf, n_sectors, loads_in_flight, threads_per_block, blocks, steps, d_sink, stream);
    case 10: return launch_v<10>(buf, n_sectors, loads_in_flight, threads_per_block, blocks, steps, d_sink, stream);
    case 11: return launch_v<11>(buf, n_sectors, loads_in_flight, threads_per_block, blocks, steps, d_sink, stream);
    case 12: return launch_v<12>(buf, n_sectors, loads_in_flight, threads_per_block, blocks, steps, d_sink, stream);
    case 13: return launch_v<13>(buf, n_sectors, loads_in_flight, threads_per_block, blocks, steps, d_sink, stream);
    default: return cudaErrorInvalidValue;
    }
}
