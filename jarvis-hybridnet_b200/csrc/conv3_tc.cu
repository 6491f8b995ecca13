// Stage 2, bf16 throughput path: the C->C 3x3x3 stride-1 convolutions of V2VNet's Res3DBlocks
// (jarvis/hybridnet/v2vnet.py:27-43; 8 of the 12 layers, 91 % of the FLOPs) as a tcgen05 implicit GEMM whose
// N dimension stacks the three x-taps:
//
//     D[r][dx*C + co] += sum_ci  X[z+dz-1][p0 + r + (dy-1)*Wp][ci] * W[co][ci][dz][dy][dx]       (dz,dy: 9 MMA groups)
//     out[p0 + r][co]  = D[r-1][0*C + co] + D[r][1*C + co] + D[r+1][2*C + co]                    (epilogue)
//
// Why: with the positions on M (128) and only C = 48 output channels on N, a tap-per-MMA formulation reads
// 4 KB of A from shared memory for a 24-cycle MMA and is bound by the 128 B/clk shared-memory port (45.5
// cycles per MMA measured, profiles/r01_umma_microbench_cycles_per_mma.txt).  Stacking dx on N makes one
// MMA 128 x 144 x 16: 72 tensor cycles for 4 KB of A + 4.5 KB of B = 68 port cycles, i.e. tensor-bound, and
// a tile needs 27 MMAs instead of 81.  The x-shift moves to the epilogue where it is a one-lane warp shuffle
// of the fp32 accumulator rows (TMEM lane == GEMM row == flat padded position), plus a two-row exchange
// through shared memory at the warp-quadrant boundaries.  Rows 0 and 127 of a tile have no left / right
// neighbour, so tiles advance by 126 positions.
//
// Tiles are ordered z-fastest and each CTA walks a contiguous tile range, so the three input planes of a tile
// are a sliding window over a ring of plane boxes in shared memory: after the first tile of a column every
// tile fetches ONE new plane box (KC contiguous TMA bulk copies) instead of three.
//
// Roles (one persistent CTA per SM):
//     warp 0      TMA producer (one elected lane): weights once (resident, 124 KB for C = 48), plane boxes
//     warps 1-2   MMA issuers (one elected lane each, alternate tiles; warp 1 also allocates TMEM)
//     warps 3-    epilogue: two warps per TMEM lane quadrant, each owning half of the output channels:
//                 tcgen05.ld -> shuffle-combine -> bias -> pad mask -> bf16 store, and the per-channel
//                 sum / sum-of-squares of the following InstanceNorm accumulated in registers across tiles
//                 (one atomicAdd per channel and warp when the CTA's range leaves a sample).
#include <atomic>
#include <cstdlib>

#include "tc_ptx.cuh"
#include "v2v.cuh"

namespace jhn {

constexpr int C3_VALID = TILE_M - 2;                   // output rows per tile
constexpr int C3_MAX_SLOTS = 8;
#ifndef C3_EW48
#define C3_EW48 3                                        // epilogue warps per quadrant for the 48-channel layers
#endif

struct C3Launch {
    const uint4 *in;                                   // BP bf16 input  [B][KC][D+2][(D+2)^2] 16-byte voxels
    const __nv_bfloat16 *w;                            // [dz*3+dy][KC][dx*NOUT+co][8] bf16
    const float *bias;                                 // [NOUT]
    uint4 *out;                                        // BP bf16 output [B][KC][D+2][(D+2)^2]
    StatPart stats;                                    // per-CTA partial sums of the following InstanceNorm (part may be null)
    int B, D, NT, total_tiles;
    int NS, PB;                                        // ring slots, positions per plane box
    int add_bias;                                      // 0 when an InstanceNorm follows (it cancels the bias exactly)
    int l2_hint;                                       // bit 0: plane boxes loaded evict-first; bit 1: output stored evict-last
};

// one 8-column piece of the three dx groups, loaded and waited for in ONE asm statement so that no consumer of
// the destination registers can be scheduled ahead of tcgen05.wait::ld
__device__ __forceinline__ void c3_ld8x3(uint32_t t0, uint32_t t1, uint32_t t2, uint32_t *a, uint32_t *b, uint32_t *c)
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%24];\n"
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%8,%9,%10,%11,%12,%13,%14,%15}, [%25];\n"
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%16,%17,%18,%19,%20,%21,%22,%23}, [%26];\n"
        "tcgen05.wait::ld.sync.aligned;\n"
        : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]), "=r"(a[4]), "=r"(a[5]), "=r"(a[6]), "=r"(a[7]),
          "=r"(b[0]), "=r"(b[1]), "=r"(b[2]), "=r"(b[3]), "=r"(b[4]), "=r"(b[5]), "=r"(b[6]), "=r"(b[7]),
          "=r"(c[0]), "=r"(c[1]), "=r"(c[2]), "=r"(c[3]), "=r"(c[4]), "=r"(c[5]), "=r"(c[6]), "=r"(c[7])
        : "r"(t0), "r"(t1), "r"(t2)
        : "memory");
}

// Two builds of the 480-thread configuration (NOUT <= 48, EW = 3).  MAXR = 128 is the fastest (0.75 ms per 32 frame sets for the
// six layers).  MAXR = 120 (0.79 ms) leaves 1024 registers free in every scheduler partition of the SM: room for one warp of
// the 32-register host-memory pull kernel (jhn_pull_heatmap_boxes) per partition.  With 128 the pull CTAs and this kernel's CTAs
// keep each other off an SM and a forward that overlaps a transfer runs the 3x3x3 layers in two waves; the launcher picks the
// 120-register build while the calling thread has announced an overlapping transfer (jhn_set_transfer_overlap; run 60:
// resident 10.32 k vs 10.19 k frame-sets/s, end to end 6.9 k vs 7.65 k).  A thread bound cannot express 120 (544 threads =
// 5 warps per partition -> 96), hence __maxnreg__.
// EW = epilogue warps per TMEM lane quadrant; each owns NOUT / EW output channels
template <int NOUT, int EW, int MAXR>
__global__ void __maxnreg__(MAXR)
tc_conv3_kernel(const C3Launch L)
{
    constexpr int C3_THREADS = 96 + 128 * EW;
    constexpr int KC = NOUT / 8;                       // 8-channel chunks of the input (Cin == Cout == NOUT)
    constexpr int N3 = 3 * NOUT;                       // MMA N: three x-taps stacked
    constexpr int CW = NOUT / EW;                      // output channels per epilogue warp
    constexpr int W_BYTES = 9 * KC * N3 * 16;
    static_assert(NOUT % 16 == 0 && N3 <= 256, "stacked N must be a legal tcgen05 N");
#ifndef C3_ACC_BUFS
#define C3_ACC_BUFS 2
#endif
    // accumulator buffers in TMEM (512 columns): three when they fit (N3 <= 160), so that the MMAs run up to two tiles ahead
    // of the epilogue; else two
    constexpr int ACC_STRIDE = (N3 + 31) / 32 * 32;
    constexpr int NB = 2;                              // one accumulator buffer per issuing thread
    constexpr int ACC_COLS = NB == 3 ? ACC_STRIDE : 256;
    static_assert(NOUT % (8 * EW) == 0, "channels per epilogue warp must be whole 8-channel chunks");

    extern __shared__ __align__(1024) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int D = L.D, Wp = D + 2, PP = Wp * Wp;
    const int slot_bytes = KC * L.PB * 16;

    uint8_t *w_smem = smem;
    uint8_t *ring = smem + W_BYTES;
    float *xch = reinterpret_cast<float *>(ring + (size_t)L.NS * slot_bytes);     // [2][4][2][NOUT]
    float *bias_s = xch + 2 * 4 * 2 * NOUT;
    uint64_t *bars = reinterpret_cast<uint64_t *>(bias_s + NOUT);
    uint64_t *full = bars, *empty = bars + C3_MAX_SLOTS, *tfull = bars + 2 * C3_MAX_SLOTS, *tempty = tfull + 3, *wbar = tempty + 3;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(wbar + 1);

    const int t_begin = (int)((long long)L.total_tiles * blockIdx.x / gridDim.x);
    const int t_end = (int)((long long)L.total_tiles * (blockIdx.x + 1) / gridDim.x);

    if (threadIdx.x == 0) {
        for (int i = 0; i < L.NS; ++i) { mbar_init(smem_u32(full + i), 1); mbar_init(smem_u32(empty + i), 3); }
        for (int i = 0; i < NB; ++i) { mbar_init(smem_u32(tfull + i), 1); mbar_init(smem_u32(tempty + i), 4 * EW); }
        mbar_init(smem_u32(wbar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < NOUT; i += C3_THREADS) bias_s[i] = L.bias[i];
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // =================================== TMA producer ===================================
        if (elect_one()) {
            mbar_expect_tx(smem_u32(wbar), (uint32_t)W_BYTES);
            for (int off = 0; off < W_BYTES; off += 32768) {
                const int n = min(32768, W_BYTES - off);
                bulk_load(smem_u32(w_smem + off), reinterpret_cast<const uint8_t *>(L.w) + off, n, smem_u32(wbar));
            }
            int slot = 0; uint32_t phase = 0;
            int z = t_begin % D, u = t_begin / D;
            const uint64_t pol_in = l2_policy_evict_first();
            const bool hint_in = (L.l2_hint & 1) != 0;
            for (int t = t_begin; t < t_end; ++t) {
                const int pt = u % L.NT, b = u / L.NT;
                const bool fresh = (t == t_begin) || (z == 0);
                const int start = C3_VALID * pt;                               // box = positions [start, start + PB)
                const int npos = min(L.PB, PP - start);
                const uint32_t run = (uint32_t)npos * 16;
                for (int dz = fresh ? 0 : 2; dz < 3; ++dz) {
                    mbar_wait(smem_u32(empty + slot), phase ^ 1);
                    const uint32_t fb = smem_u32(full + slot);
#ifdef C3_DBG_NO_TMA
                    mbar_arrive(fb);
                    if (++slot == L.NS) { slot = 0; phase ^= 1; }
                    continue;
#endif
                    mbar_expect_tx(fb, run * KC);
                    const uint4 *src = L.in + ((size_t)b * KC * Wp + (z + dz)) * PP + start;
                    uint8_t *dst = ring + (size_t)slot * slot_bytes;
#pragma unroll
                    for (int j = 0; j < KC; ++j) {
                        if (hint_in) bulk_load_hint(smem_u32(dst + (size_t)j * L.PB * 16), src + (size_t)j * Wp * PP, run, fb, pol_in);
                        else bulk_load(smem_u32(dst + (size_t)j * L.PB * 16), src + (size_t)j * Wp * PP, run, fb);
                    }
                    if (++slot == L.NS) { slot = 0; phase ^= 1; }
                }
                if (++z == D) { z = 0; ++u; }
            }
        }
    } else if (warp <= 2) {
        // =================================== MMA issuers =====================================
        // TWO issuing threads (warps 1 and 2, on different schedulers) take alternate tiles of the CTA's range, each into the
        // accumulator buffer of its own.  One thread needs ~100 cycles of instruction time per N = 144 MMA (descriptor
        // arithmetic, R2UR, barrier waits; the tensor pipe needs 72 and tools/umma_bench.cu reaches that only with a bare
        // issue loop), and the pipe queues too few MMAs to hide a prep block: with two threads each has two tile times per tile.
        // MMAs of different tiles may interleave in the pipe: they accumulate into different TMEM buffers.
        // A plane box is read by up to three consecutive tiles (dz = 2, 1, 0), i.e. by both threads: its empty barrier
        // counts three arrivals, one tcgen05.commit per reading tile; at the ends of a column segment, where a plane has
        // fewer readers, the last reader adds the missing arrivals.
        if (elect_one()) {
            const int me = warp - 1;
            const uint32_t idesc = umma_idesc_bf16(N3);
            const uint64_t hi_c = (uint64_t)(8u | (1u << 14)) << 32;           // SBO = 128 B, descriptor version 1
            const uint32_t lbo_a = (uint32_t)L.PB << 16, lbo_b = (uint32_t)N3 << 16;   // K-chunk strides, 16-byte units
            const uint32_t w_units = smem_u32(w_smem) >> 4, ring_units = smem_u32(ring) >> 4;
            const uint32_t slot_units = (uint32_t)slot_bytes >> 4;
            mbar_wait(smem_u32(wbar), 0);
            // ring position of the oldest plane box of the current tile as (slot, wrap parity): no integer division on the issue path
            int s0 = 0; uint32_t w0 = 0;
            uint32_t aphase = 0;
            const uint32_t d_tmem = tmem_base + (uint32_t)(me * ACC_COLS);
            int z = t_begin % D;
            int seg_i = 0;                                                     // index of the tile inside its column segment
            for (int t = t_begin; t < t_end; ++t) {
                const bool fresh = (t == t_begin) || (z == 0);
                const bool last = (t + 1 == t_end) || (z + 1 == D);           // the column segment ends with this tile
                if (t != t_begin) { s0 += fresh ? 3 : 1; if (s0 >= L.NS) { s0 -= L.NS; w0 ^= 1u; } }
                if (fresh) seg_i = 0;
                if (((t - t_begin) & 1) == me) {
                    mbar_wait(smem_u32(tempty + me), aphase ^ 1);
                    tc_fence_after();
                    aphase ^= 1;
                    uint32_t acc = 0;
                    // all three planes are waited for up front and the 27 MMAs of the tile are one straight-line block: the
                    // descriptor arithmetic of the whole tile runs while the OTHER thread's MMAs execute, and nothing but
                    // UTCHMMA / R2UR sits between this tile's MMAs (a prep block per plane left the pipe idle three times a tile)
                    int slot[3];
#pragma unroll
                    for (int dz = 0; dz < 3; ++dz) {
                        slot[dz] = s0 + dz; uint32_t wrap = w0;
                        if (slot[dz] >= L.NS) { slot[dz] -= L.NS; wrap ^= 1u; }
                        // every plane is waited for by the thread that reads it: the other thread observed the planes this tile
                        // shares with its predecessor, this one may not have (second tile of a segment)
                        mbar_wait(smem_u32(full + slot[dz]), wrap);
                    }
                    tc_fence_after();
#pragma unroll
                    for (int dz = 0; dz < 3; ++dz) {
                        const uint32_t a0 = ring_units + (uint32_t)slot[dz] * slot_units;
                        const uint32_t b0 = w_units + (uint32_t)(dz * 3 * KC * N3);
#pragma unroll
                        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                            for (int kc = 0; kc < KC; kc += 2) {
                                const uint32_t a_lo = lbo_a | (a0 + (uint32_t)(dy * Wp + kc * L.PB));
                                const uint32_t b_lo = lbo_b | (b0 + (uint32_t)((dy * KC + kc) * N3));
#ifndef C3_DBG_NO_MMA
                                tc_mma_bf16(d_tmem, hi_c | (uint64_t)a_lo, hi_c | (uint64_t)b_lo, idesc, acc);
#endif
                                acc = 1;
                            }
                        const uint32_t eb = smem_u32(empty + slot[dz]);
                        tc_commit(eb);                                         // this tile's read of the plane box
                        if (dz == 0 || last) {
                            // last reader of plane j = seg_i + dz of the segment: readers are the tiles max(0, j-2) .. min(n-1, j);
                            // n is not known before the segment ends, but "last" says whether this tile is n - 1
                            const int j = seg_i + dz;
                            const int first_reader = j - 2 > 0 ? j - 2 : 0;
                            const int readers = seg_i - first_reader + 1;      // this tile is the last reader
                            for (int m = readers; m < 3; ++m) mbar_arrive(eb);
                        }
                    }
                    tc_commit(smem_u32(tfull + me));
                }
                ++seg_i;
                if (++z == D) z = 0;
            }
        }
    } else {
        // =================================== epilogue warps =================================
        const int q = warp & 3;                                                // TMEM lane quadrant
        const int part = (warp - 3) >> 2;                                      // which slice of the channels
        const int row = q * 32 + lane;
        const int c_base = part * CW;
        float s1[CW], s2[CW], bias_r[CW];
#pragma unroll
        for (int i = 0; i < CW; ++i) { s1[i] = 0.f; s2[i] = 0.f; bias_r[i] = L.add_bias ? bias_s[c_base + i] : 0.f; }
        int stat_b = -1;
        auto flush = [&](int b) {
            const int slot = ((int)blockIdx.x + b) * 4 + q;
            float2 *dst = reinterpret_cast<float2 *>(L.stats.part) + (size_t)slot * NOUT + c_base;
#pragma unroll
            for (int i = 0; i < CW; ++i) {
                float a = s1[i], c = s2[i];
#pragma unroll
                for (int s = 16; s > 0; s >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, s); c += __shfl_xor_sync(0xffffffffu, c, s); }
                if (lane == 0) dst[i] = make_float2(a, c);
                s1[i] = 0.f; s2[i] = 0.f;
            }
        };
        int ab = 0; uint32_t aphase = 0;
        // tile coordinates advance z-fastest; everything that depends on (b, pt) only is refreshed on a column change
        int z = t_begin % D, pt = (t_begin / D) % L.NT, b = (t_begin / D) / L.NT;
        bool rowok = false, valid = false;
        uint4 *optr = nullptr;                                                 // output voxel of this row in plane z, chunk c_base/8
        const uint64_t pol_out = l2_policy_evict_last();
        const bool hint_out = (L.l2_hint & 2) != 0;
        bool column_changed = true;
        const size_t plane_stride = (size_t)PP, chunk_stride = (size_t)Wp * PP;
        const uint32_t taddr0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c_base;
        for (int t = t_begin; t < t_end; ++t) {
            if (column_changed) {
                if (L.stats.part && b != stat_b) {
                    if (stat_b >= 0) flush(stat_b);
                    stat_b = b;
                }
                const int p = Wp + C3_VALID * pt + row;
                const int yp = p / Wp, xp = p - yp * Wp;
                rowok = row >= 1 && row <= C3_VALID && p < PP;
                valid = rowok && xp >= 1 && xp <= D && yp >= 1 && yp <= D;
                optr = L.out + (((size_t)b * KC + (c_base >> 3)) * Wp + (z + 1)) * plane_stride + p;
                column_changed = false;
            }
            float *xw = xch + (size_t)(t & 1) * 4 * 2 * NOUT;
            mbar_wait(smem_u32(tfull + ab), aphase);
            tc_fence_after();
#ifdef C3_DBG_NO_EPI
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(tempty + ab));
            optr += plane_stride;
            if (++ab == NB) { ab = 0; aphase ^= 1; }
            if (++z == D) { z = 0; column_changed = true; if (++pt == L.NT) { pt = 0; ++b; } }
            continue;
#endif
            float o[CW];
            const uint32_t taddr = taddr0 + (uint32_t)(ab * ACC_COLS);
#pragma unroll
            for (int pc = 0; pc < CW / 8; ++pc) {
                uint32_t d0[8], d1[8], d2[8];
                c3_ld8x3(taddr + pc * 8, taddr + NOUT + pc * 8, taddr + 2 * NOUT + pc * 8, d0, d1, d2);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float up = __shfl_up_sync(0xffffffffu, __uint_as_float(d0[i]), 1);
                    const float dn = __shfl_down_sync(0xffffffffu, __uint_as_float(d2[i]), 1);
                    float v = __uint_as_float(d1[i]);
                    if (lane > 0) v += up;
                    if (lane < 31) v += dn;
                    o[pc * 8 + i] = v;
                }
                if (lane == 31) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) xw[(q * 2 + 0) * NOUT + c_base + pc * 8 + i] = __uint_as_float(d0[i]);
                }
                if (lane == 0) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) xw[(q * 2 + 1) * NOUT + c_base + pc * 8 + i] = __uint_as_float(d2[i]);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(tempty + ab));                 // accumulator buffer free again
            // quadrant-boundary rows are in xch: only the four warps of one channel slice exchange rows, so each
            // slice synchronises on its own named barrier (128 threads) instead of all 128 * EW epilogue threads
            asm volatile("bar.sync %0, 128;" ::"r"(1 + part) : "memory");
            if (lane == 0 && q > 0) {
#pragma unroll
                for (int i = 0; i < CW; ++i) o[i] += xw[((q - 1) * 2 + 0) * NOUT + c_base + i];
            }
            if (lane == 31 && q < 3) {
#pragma unroll
                for (int i = 0; i < CW; ++i) o[i] += xw[((q + 1) * 2 + 1) * NOUT + c_base + i];
            }
#pragma unroll
            for (int i = 0; i < CW; ++i) {
                const float v = valid ? o[i] + bias_r[i] : 0.f;
                o[i] = v;
                s1[i] += v;
                s2[i] = fmaf(v, v, s2[i]);
            }
            if (rowok) {
#pragma unroll
                for (int j = 0; j < CW / 8; ++j) {
                    uint32_t pk[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        __nv_bfloat162 h2 = __floats2bfloat162_rn(o[8 * j + 2 * i], o[8 * j + 2 * i + 1]);
                        pk[i] = *reinterpret_cast<uint32_t *>(&h2);
                    }
                    if (hint_out) st_global_hint(optr + (size_t)j * chunk_stride, make_uint4(pk[0], pk[1], pk[2], pk[3]), pol_out);
                    else optr[(size_t)j * chunk_stride] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                }
            }
            optr += plane_stride;
            if (++ab == NB) { ab = 0; aphase ^= 1; }
            if (++z == D) {
                z = 0; column_changed = true;
                if (++pt == L.NT) { pt = 0; ++b; }
            }
        }
        if (L.stats.part && stat_b >= 0) flush(stat_b);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// [dz*3+dy][KC][dx*NOUT+co][8] bf16 from PyTorch Conv3d fp32 [cout][cin][3][3][3], zero padded
__global__ void c3_pack_weights_kernel(const float *__restrict__ src, __nv_bfloat16 *__restrict__ dst, int cout, int cin, int NOUT)
{
    const int KC = NOUT / 8, N3 = 3 * NOUT;
    const int n = 9 * KC * N3 * 8;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const int i = e & 7; int r = e >> 3;
    const int nn = r % N3; r /= N3;
    const int j = r % KC;
    const int g = r / KC;                                                      // dz*3+dy
    const int dx = nn / NOUT, co = nn - dx * NOUT, ci = j * 8 + i;
    float v = 0.f;
    if (co < cout && ci < cin) v = src[((size_t)co * cin + ci) * 27 + g * 3 + dx];
    dst[e] = __float2bfloat16_rn(v);
}

static size_t c3_tail_bytes(int NOUT) { return (size_t)(2 * 4 * 2 * NOUT + NOUT) * 4 + (2 * C3_MAX_SLOTS + 7) * 8 + 16; }

size_t c3_weight_bytes(int NOUT) { return (size_t)9 * (NOUT / 8) * 3 * NOUT * 16; }

// Can the stacked kernel run a C->C 3x3x3 layer of padded width NOUT on grid side D?  Fills ring geometry.
bool c3_plan(int NOUT, int D, int max_smem, int *NS, int *PB)
{
    if (NOUT % 16 != 0 || NOUT < 16 || NOUT > 80) return false;
    const int Wp = D + 2;
    const int pb = TILE_M + 2 * Wp;
    const size_t slot = (size_t)(NOUT / 8) * pb * 16;
    const long long room = (long long)max_smem - (long long)c3_weight_bytes(NOUT) - (long long)c3_tail_bytes(NOUT) - 1024;
    if (room < 0) return false;
    int ns = (int)(room / (long long)slot);
    if (ns > C3_MAX_SLOTS) ns = C3_MAX_SLOTS;
    if (ns < 4) return false;                                                  // three planes in use + one in flight
    *NS = ns; *PB = pb;
    return true;
}

int c3_pack(const float *src, __nv_bfloat16 *dst, int cout, int cin, int NOUT, cudaStream_t st)
{
    const int n = 9 * (NOUT / 8) * 3 * NOUT * 8;
    JHN_LAUNCH("c3_pack_weights_kernel", st, c3_pack_weights_kernel<<<cdiv(n, 256), 256, 0, st>>>(src, dst, cout, cin, NOUT));
    return JHN_OK;
}

static thread_local int t_transfer_overlap = 0;
void c3_set_transfer_overlap(int on) { t_transfer_overlap = on ? 1 : 0; }

template <int NOUT, int EW, int MAXR>
static int c3_launch_t(const C3Launch &L, int grid, size_t smem, cudaStream_t st)
{
    // the attribute is per (function, device): one bit per device, set on first use, safe from any thread
    static std::atomic<unsigned long long> configured{0ull};
    int dev = 0;
    JHN_CUDA(cudaGetDevice(&dev));
    const unsigned long long bit = 1ull << (dev & 63);
    if (!(configured.load(std::memory_order_acquire) & bit)) {
        int max_smem = 0;
        JHN_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
        JHN_CUDA(cudaFuncSetAttribute(tc_conv3_kernel<NOUT, EW, MAXR>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        configured.fetch_or(bit, std::memory_order_release);
    }
    JHN_LAUNCH("tc_conv3_stacked", st, tc_conv3_kernel<NOUT, EW, MAXR><<<grid, 96 + 128 * EW, smem, st>>>(L));
    return JHN_OK;
}

int c3_launch(int NOUT, const void *in, const __nv_bfloat16 *w, const float *bias, void *out, StatPart *stats, int B, int D,
              int NS, int PB, int sms, int add_bias, cudaStream_t st)
{
    C3Launch L;
    L.in = (const uint4 *)in; L.w = w; L.bias = bias; L.out = (uint4 *)out; L.B = B; L.D = D;
    const int Wp = D + 2;
    L.NT = cdiv((long long)(D - 1) * Wp + D, C3_VALID);
    L.total_tiles = B * L.NT * D;
    L.NS = NS; L.PB = PB; L.add_bias = add_bias;
    // output stored with L2 evict-last priority: the normalisation pass that follows finds a little more of it in L2
    // (11 passes 0.689 -> 0.653 ms per 32 frame sets, run 43; evict-first on the plane boxes changed nothing)
    static const int l2_hint = [] { const char *e = getenv("JHN_C3_L2HINT"); return e ? atoi(e) : 2; }();
    L.l2_hint = l2_hint;
    const size_t smem = c3_weight_bytes(NOUT) + (size_t)NS * (NOUT / 8) * PB * 16 + c3_tail_bytes(NOUT);
    const int grid = L.total_tiles < sms ? L.total_tiles : sms;
    if (stats) { stats->grid = grid; stats->Tb = L.NT * D; stats->T = L.total_tiles; L.stats = *stats; }
    else L.stats = StatPart{nullptr, 0, 0, 0};
    switch (NOUT) {
    case 16: return c3_launch_t<16, 2, 168>(L, grid, smem, st);
    case 32: return c3_launch_t<32, 2, 168>(L, grid, smem, st);
    case 48: return t_transfer_overlap ? c3_launch_t<48, C3_EW48, 120>(L, grid, smem, st) : c3_launch_t<48, C3_EW48, 128>(L, grid, smem, st);
    case 64: return c3_launch_t<64, 2, 168>(L, grid, smem, st);
    case 80: return c3_launch_t<80, 2, 168>(L, grid, smem, st);
    }
    return fail(JHN_ERR_SHAPE, "stacked 3x3x3 kernel: unsupported channel width %d", NOUT);
}

}  // namespace jhn
