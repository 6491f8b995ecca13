// Stage 2, bf16 throughput path: V2VNet (jarvis/hybridnet/v2vnet.py:12-102) as implicit GEMMs on the
// 5th-gen tensor cores (tcgen05.mma, accumulators in TMEM, operands staged by TMA).
//
// Data layout ("BP", blocked + padded), chosen so that EVERY filter tap of a 3-D convolution is a plain
// 16-byte-granular shift of one shared-memory tile:
//     act[b][j][zp][pp][8]   bf16,  j = channel/8,  zp in [0,D+2),  pp = yp*(D+2)+xp in [0,(D+2)^2)
// i.e. for each 8-channel chunk the zero-padded volume is a flat array of 16-byte voxels.  A GEMM row
// block is 128 consecutive padded positions of one z-plane; the input window of tap (dz,dy,dx) is the
// same flat array shifted by dy*(D+2)+dx positions in plane z+dz.  A box [KC chunks][PB positions]
// (PB = 128 + halo; one contiguous TMA bulk copy per chunk) lands in smem as [chunk][position][8] which IS the canonical no-swizzle K-major UMMA
// operand layout (8 rows x 16 B core matrices, SBO = 128 B between row groups, LBO = PB*16 B between the
// two K-chunks of one K=16 MMA) for any start row -> the 9 in-plane taps are 9 descriptors into one tile,
// and the tile is fetched from L2 once per plane instead of once per tap.  Pad positions produce garbage
// rows that the epilogue replaces by zeros, which keeps the zero border of the output tensor intact.
// Stride-2 convolutions read a parity-split ("PS") copy [b][8 sub-volumes][j][zp][pp][8] of their input in
// which every tap is again a unit-stride shift; ConvTranspose3d(k2,s2) is 8 independent 1x1x1 GEMMs whose
// epilogue scatters to the 8 output parities.
//
// Kernel: persistent, warp-specialised, one CTA per SM.
//     warps 0-3  TMA producers (one elected lane each): cp.async.bulk of activation boxes (+ weight slabs when not
//                resident) into a ring of shared-memory slots, completion on mbarriers
//     warps 4-7  MMA issuers (one elected lane each; warp 4 also allocates TMEM): tcgen05.mma M=128, N=Cout_pad, K=16,
//                descriptors formed from a per-tile offset list precomputed in shared memory.
//                Producer r, ring share r, issuer r and accumulator buffer r form pipeline r; tile i of the CTA goes to
//                pipeline i % npipe.  One producer / issuer thread per CTA was the bound of the k3 s2 front convolution:
//                each spent ~100 dependent instructions per stage (12 stages per tile), 70-80 % of its cycles issuing
//                (profiles/r02_tc_conv_front_ncu.txt).  npipe = 4 when the ring has >= 8 slots and 4 accumulator buffers
//                fit TMEM, else 2, else 1.
//     warps 8-15 epilogue, two per TMEM lane quadrant on alternate 16-column chunks: tcgen05.ld -> bias -> pad mask -> bf16 store and
//                per-channel sum / sum-of-squares for the following InstanceNorm (fused statistics)
// Weights of the C->C 3x3x3 layers (124 KB bf16) stay resident in shared memory for the CTA's lifetime.
// The work of a layer is a "stage program" (TcProgram) built on the host: per tile a list of TMA boxes and,
// per box, the taps (smem row offset, weight index) to issue against it.
#include <cstdlib>
#include <cstring>

#include "tc_ptx.cuh"
#include "v2v.cuh"

namespace jhn {

// ------------------------------------------------------------------------------------------------
// program description shared by host and device
// ------------------------------------------------------------------------------------------------
enum { EPI_RAW = 0, EPI_CONVT = 1, EPI_HEAD = 2 };
constexpr int MAX_STAGES = 12, MAX_TAPS = 9;

struct TcTap { int16_t aoff, widx; };                 // row offset inside the box; weight tap index
struct TcStage {
    int16_t chunk0;                                   // first input chunk of the box (per sample)
    int16_t dz;                                       // plane offset of the box relative to output z
    int32_t pos_off;                                  // first position of the box relative to tile p0
    int16_t ntaps, wtap0;                             // taps issued against this box; first weight tap
    TcTap taps[MAX_TAPS];
};
struct TcProgram {
    int nstages, KC, NOUT, PB, resident, ntaps_total, epi, tile_taps;
    int stage_bytes_a, stage_bytes, nslots;           // smem ring geometry (host computed)
    int w_bytes;                                      // resident weight bytes (0 if streamed)
    int KC_load;                                      // input chunks that carry real channels: the all-padding chunks
                                                      // above it are zeroed once in shared memory and never fetched
    TcStage st[MAX_STAGES];
};

struct TcLaunch {
    const void *in;                                   // BP / PS bf16 input tensor
    const __nv_bfloat16 *w;                           // [tap][KC][NOUT][8] bf16
    const float *bias;                                // [NOUT] fp32 (zero padded)
    void *out;                                        // BP bf16 (RAW/CONVT) or NCDHW fp32 (HEAD)
    StatPart stats;                                   // per-CTA partial (sum, sumsq) of the following InstanceNorm (part may be null)
    int B, D;                                         // samples, grid side of the GEMM-row positions
    int CJ_in;                                        // chunks per sample in the input tensor (TMA dim 3)
    int CJ_out;                                       // chunks per sample in the output tensor
    int NT;                                           // tiles per z-plane
    int total_tiles;                                  // B * D * NT * tile_taps
    int Kout;                                         // real output channels (HEAD)
};

// ------------------------------------------------------------------------------------------------
// the implicit-GEMM kernel
// ------------------------------------------------------------------------------------------------
constexpr int TC_PIPES = 4;                          // producer / issuer warp pairs, each with its own ring share and accumulator buffer
constexpr int TC_THREADS = 32 * (2 * TC_PIPES + 8);  // TC_PIPES TMA warps, TC_PIPES MMA warps, 8 epilogue warps (two per TMEM lane quadrant)
constexpr int TC_EPI_WARPS = 8;
constexpr int EPI_TILE_FLOATS = 32 * 17;
constexpr int TC_MAX_SLOTS = 16;                       // ring depth: bytes in flight per SM = slots x box; 8 slots of the
                                                      // 10 KB front-conv boxes cover only ~2/3 of the L2 latency-bandwidth product
constexpr int MAX_DESC = 1024;                         // descriptor pairs: (MMAs per tile) x (ring slots)
constexpr int TC_SMEM_TAIL = TC_EPI_WARPS * EPI_TILE_FLOATS * 4 + 128 * 4 + (2 * TC_MAX_SLOTS + 2 * 4 + 10) * 8 + MAX_DESC * 8 + 1024;

struct TileCoord { int b, z, pt, tap; };
__device__ __forceinline__ TileCoord decode_tile(int t, const TcLaunch &L, int tile_taps)
{
    TileCoord c;
    // z fastest: consecutive tiles of a CTA's range share their in-plane position, so the input planes a tile
    // re-reads for z + 1 were fetched one tile ago and are still in L2 (with the in-plane index fastest the reuse
    // distance was a whole plane of tiles, more than a CTA's share of L2)
    c.tap = t % tile_taps; t /= tile_taps;
    c.z = t % L.D; t /= L.D;
    c.pt = t % L.NT;
    c.b = t / L.NT;
    return c;
}

__global__ void __launch_bounds__(TC_THREADS, 1)
tc_conv_kernel(const __grid_constant__ TcProgram P, const TcLaunch L)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int Wp = L.D + 2, PP = Wp * Wp;
    const int tap_bytes = P.KC * P.NOUT * 16;

    // ---- carve shared memory -----------------------------------------------------------------
    uint8_t *w_smem = smem;                                            // resident weights
    uint8_t *ring = smem + P.w_bytes;                                  // nslots x stage_bytes
    float *epi_tiles = reinterpret_cast<float *>(ring + (size_t)P.nslots * P.stage_bytes);
    float *bias_s = epi_tiles + TC_EPI_WARPS * EPI_TILE_FLOATS;
    uint64_t *bars = reinterpret_cast<uint64_t *>(bias_s + 128);
    uint64_t *full = bars, *empty = bars + TC_MAX_SLOTS, *tfull = bars + 2 * TC_MAX_SLOTS, *tempty = tfull + TC_PIPES, *wbar = tfull + 2 * TC_PIPES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(wbar + 1);
    int *stage_first = reinterpret_cast<int *>(wbar + 2);             // [MAX_STAGES + 1]
    uint2 *desc_list = reinterpret_cast<uint2 *>(wbar + 10);          // [nslots][MMAs per tile] (A desc lo, B desc lo)

    // contiguous tile range of this CTA
    const int t_begin = (int)((long long)L.total_tiles * blockIdx.x / gridDim.x);
    const int t_end = (int)((long long)L.total_tiles * (blockIdx.x + 1) / gridDim.x);
    // TMEM: 2 tile buffers x 2 interleaved accumulators (the K loop of a tile alternates between two
    // independent accumulation chains; the epilogue adds them) x NOUT fp32 columns
    // pipelines: each needs >= 2 ring slots and an accumulator buffer of 2 x NOUT columns (two interleaved accumulation chains)
    const int npipe = (P.nslots >= 8 && 8 * P.NOUT <= 512) ? 4 : (P.nslots >= 4 ? 2 : 1);
    const int nbuf = npipe > 2 ? npipe : 2;                           // accumulator buffers
    const uint32_t tmem_need = (uint32_t)(nbuf * 2 * P.NOUT);
    const uint32_t tmem_cols = tmem_need <= 32 ? 32 : tmem_need <= 64 ? 64 : tmem_need <= 128 ? 128 : tmem_need <= 256 ? 256 : 512;

    bool dual = true;                                 // second accumulator is written in every stage?
    for (int s = 0; s < P.nstages; ++s) dual = dual && (P.st[s].ntaps * (P.KC / 2) >= 2);
    if (threadIdx.x == 0) {
        for (int i = 0; i < P.nslots; ++i) { mbar_init(smem_u32(full + i), 1); mbar_init(smem_u32(empty + i), 1); }
        for (int i = 0; i < TC_PIPES; ++i) { mbar_init(smem_u32(tfull + i), 1); mbar_init(smem_u32(tempty + i), TC_EPI_WARPS); }
        mbar_init(smem_u32(wbar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 128; i += TC_THREADS) bias_s[i] = i < P.NOUT ? L.bias[i] : 0.f;
    if (P.KC_load < P.KC) {
        // channel-padding chunks of the A operand: zero in every ring slot, for the whole kernel (TMA never writes them)
        const int per_slot = (P.KC - P.KC_load) * P.PB;
        for (int i = threadIdx.x; i < per_slot * P.nslots; i += TC_THREADS) {
            const int slot = i / per_slot, r = i - slot * per_slot;
            reinterpret_cast<uint4 *>(ring + (size_t)slot * P.stage_bytes)[(size_t)P.KC_load * P.PB + r] = make_uint4(0, 0, 0, 0);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == TC_PIPES) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < TC_PIPES) {
        // =================================== TMA producers ==================================
        if (elect_one()) {
            if (P.resident && warp == 0) {
                mbar_expect_tx(smem_u32(wbar), (uint32_t)P.w_bytes);
                for (int off = 0; off < P.w_bytes; off += 32768) {
                    const int n = min(32768, P.w_bytes - off);
                    bulk_load(smem_u32(w_smem + off), reinterpret_cast<const uint8_t *>(L.w) + off, n, smem_u32(wbar));
                }
            }
            // two independent pipelines (producer r -> ring half r -> issuer r -> accumulator buffer r) on alternate tiles: each
            // is a strictly sequential single-producer / single-consumer ring, so the one-bit phase parity of its mbarriers
            // stays unambiguous.  Rings of fewer than 4 slots (the streamed 96-channel layers) run as one pipeline.
            const int nsl = P.nslots / npipe, slot0 = warp * nsl;
            int slot = slot0; uint32_t phase = 0;
            for (int t = t_begin + warp; warp < npipe && t < t_end; t += npipe) {
                const TileCoord c = decode_tile(t, L, P.tile_taps);
                const int p0 = Wp + 1 + c.pt * TILE_M;
                for (int s = 0; s < P.nstages; ++s) {
                    const TcStage &S = P.st[s];
                    mbar_wait(smem_u32(empty + slot), phase ^ 1);
                    const uint32_t fb = smem_u32(full + slot);
                    uint8_t *dst = ring + (size_t)slot * P.stage_bytes;
                    // one contiguous run of positions per channel chunk; clipped at the end of the plane (the
                    // rows left stale only feed pad outputs, which the epilogue overwrites with zeros)
                    const int start = p0 + S.pos_off;
                    const int npos = min(P.PB, PP - start);
                    const uint32_t run = (uint32_t)npos * 16;
                    uint32_t bytes = run * (uint32_t)P.KC_load;
                    if (!P.resident) bytes += (uint32_t)(S.ntaps * tap_bytes);
                    mbar_expect_tx(fb, bytes);
                    const size_t unit0 = ((size_t)(c.b * L.CJ_in + S.chunk0) * Wp + (c.z + S.dz)) * PP + start;
                    const uint8_t *src = reinterpret_cast<const uint8_t *>(L.in) + unit0 * 16;
                    const size_t chunk_stride = (size_t)Wp * PP * 16;
                    for (int j = 0; j < P.KC_load; ++j)
                        bulk_load(smem_u32(dst + (size_t)j * P.PB * 16), src + j * chunk_stride, run, fb);
                    if (!P.resident)
                        bulk_load(smem_u32(dst + P.stage_bytes_a),
                                  reinterpret_cast<const uint8_t *>(L.w) + (size_t)S.wtap0 * tap_bytes,
                                  (uint32_t)(S.ntaps * tap_bytes), fb);
                    if (++slot == slot0 + nsl) { slot = slot0; phase ^= 1; }
                }
            }
        }
    } else if (warp < 2 * TC_PIPES) {
        // =================================== MMA issuers =====================================
        // Flatten the stage program into complete descriptor low words, one (A, B) pair per MMA and ring slot,
        // so that the issue loop is LDS.64 -> 2x R2UR -> UTCHMMA.  Descriptor words: lo = start address
        // (16-byte units) | LBO << 16, hi = SBO (= 8 units, 128 B) | version 1.
        const uint32_t lbo_a = (uint32_t)P.PB, lbo_b = (uint32_t)P.NOUT;           // in 16-byte units
        const uint32_t tap_units = (uint32_t)tap_bytes >> 4;
        // the TC_PIPES issuer warps fill the tables together (each entry written once) and meet on a named barrier: every
        // reader is then ordered after every writer by a barrier the whole group took part in
        const int iw = warp - TC_PIPES;
        int n_mma = 0;
        for (int s = 0; s < P.nstages; ++s) {
            if (iw == 0 && lane == 0) stage_first[s] = n_mma;
            n_mma += P.st[s].ntaps * (P.KC / 2);
        }
        if (iw == 0 && lane == 0) stage_first[P.nstages] = n_mma;
        {
            const uint32_t a_lo_c = lbo_a << 16, b_lo_c = lbo_b << 16;
            const uint32_t w_units = smem_u32(w_smem) >> 4, ring_units = smem_u32(ring) >> 4;
            const uint32_t slot_units = (uint32_t)P.stage_bytes >> 4, a_units = (uint32_t)P.stage_bytes_a >> 4;
            const int kpt = P.KC / 2;
            for (int i = iw * 32 + lane; i < n_mma * P.nslots; i += 32 * TC_PIPES) {
                const int slot = i / n_mma;
                int e = i - slot * n_mma, s = 0;
                while (e >= P.st[s].ntaps * kpt) { e -= P.st[s].ntaps * kpt; ++s; }
                const int k = e / kpt, kc = 2 * (e - k * kpt);
                const TcStage &S = P.st[s];
                const uint32_t wsel = P.tile_taps > 1 ? 0u : (P.resident ? (uint32_t)S.taps[k].widx : (uint32_t)k);
                const uint32_t a0 = ring_units + (uint32_t)slot * slot_units;
                const uint32_t b0 = P.resident ? w_units : a0 + a_units;
                desc_list[i] = make_uint2(a_lo_c | (a0 + (uint32_t)S.taps[k].aoff + kc * lbo_a),
                                          b_lo_c | (b0 + wsel * tap_units + kc * lbo_b));
            }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(32 * TC_PIPES) : "memory");
        if (elect_one()) {
            const uint32_t idesc = umma_idesc_bf16(P.NOUT);
            const uint64_t hi_c = (uint64_t)(8u | (1u << 14)) << 32;
            const uint32_t nout = (uint32_t)P.NOUT;
            if (P.resident) mbar_wait(smem_u32(wbar), 0);
            const int nsl = P.nslots / npipe, pipe = warp - TC_PIPES, slot0 = pipe * nsl;
            int slot = slot0; uint32_t phase = 0;
            int ab = pipe; uint32_t aphase = 0;                      // npipe > 1: issuer r owns accumulator buffer r; one pipeline: it alternates between two
            for (int t = t_begin + pipe; pipe < npipe && t < t_end; t += npipe) {
                const uint32_t tile_b = P.tile_taps > 1 ? (uint32_t)(t & (P.tile_taps - 1)) * tap_units : 0u;   // tile_taps is 1 or 8 (build_program)
                mbar_wait(smem_u32(tempty + ab), aphase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(ab * 2 * P.NOUT);
                uint32_t acc0 = 0, acc1 = 0;
                for (int s = 0; s < P.nstages; ++s) {
                    mbar_wait(smem_u32(full + slot), phase);
                    tc_fence_after();
                    const uint2 *dl = desc_list + slot * n_mma;
                    const int e1 = stage_first[s + 1];
                    int e = stage_first[s];
                    // descriptor pairs are fetched one step AHEAD of the MMAs that use them: the MMA wrapper is a compiler-level
                    // memory barrier, so a load placed after it would start only once the previous MMA has been issued and its
                    // LDS -> R2UR latency would sit between every two MMAs
                    if (dual) {
                        uint2 n0 = dl[e], n1 = dl[e + 1 < e1 ? e + 1 : e];
                        for (; e + 1 < e1; e += 2) {
                            const uint2 o0 = n0, o1 = n1;
                            if (e + 2 < e1) { n0 = dl[e + 2]; n1 = dl[e + 3 < e1 ? e + 3 : e + 2]; }
                            tc_mma_bf16(d_tmem, hi_c | (uint64_t)o0.x, hi_c | (uint64_t)(o0.y + tile_b), idesc, acc0);
                            tc_mma_bf16(d_tmem + nout, hi_c | (uint64_t)o1.x, hi_c | (uint64_t)(o1.y + tile_b), idesc, acc1);
                            acc0 = 1; acc1 = 1;
                        }
                        if (e < e1) {                                      // odd count: the last descriptor is already in n0
                            tc_mma_bf16(d_tmem, hi_c | (uint64_t)n0.x, hi_c | (uint64_t)(n0.y + tile_b), idesc, acc0);
                            acc0 = 1; ++e;
                        }
                    } else if (e < e1) {
                        uint2 n = dl[e];
                        for (; e < e1; ++e) {
                            const uint2 o = n;
                            if (e + 1 < e1) n = dl[e + 1];
                            tc_mma_bf16(d_tmem, hi_c | (uint64_t)o.x, hi_c | (uint64_t)(o.y + tile_b), idesc, acc0);
                            acc0 = 1;
                        }
                    }
                    tc_commit(smem_u32(empty + slot));               // frees the smem slot when the MMAs retire
                    if (++slot == slot0 + nsl) { slot = slot0; phase ^= 1; }
                }
                tc_commit(smem_u32(tfull + ab));                     // accumulator ready for the epilogue
                if (npipe > 1) aphase ^= 1;
                else if (++ab == 2) { ab = 0; aphase ^= 1; }
            }
        }
    } else {
        // =================================== epilogue warps =================================
        // two warps per TMEM lane quadrant (warp % 4), taking alternate 16-column chunks: the per-tile chain
        // tcgen05.ld -> pack -> store -> statistics is latency-bound, so a second warp per scheduler halves it
        const int wq = warp & 3;                                     // TMEM lane quarter this warp may read
        const int chalf = (warp - 2 * TC_PIPES) >> 2;                // which chunks: ci % 2 == chalf
        const int m = wq * 32 + lane;                                // GEMM row == position offset in the tile
        float *tile = epi_tiles + (warp - 2 * TC_PIPES) * EPI_TILE_FLOATS;
        const int col = lane & 15, which = lane >> 4;                // statistics ownership
        float acc_stat[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        int stat_b = -1;
        auto flush_stats = [&](int b) {               // this warp's slot of sample b: plain stores, summed in slot order by the consumer
            const int slot = ((int)blockIdx.x + b) * 4 + wq;
            float *dst = L.stats.part + (size_t)slot * P.NOUT * 2;
#pragma unroll
            for (int ci = 0; ci < 6; ++ci)
                if (ci * 16 < P.NOUT && (ci & 1) == chalf) { dst[(ci * 16 + col) * 2 + which] = acc_stat[ci]; acc_stat[ci] = 0.f; }
        };
        int ab = 0; uint32_t aphase = 0;
        const size_t plane16 = (size_t)PP;                           // 16-byte units per z-plane
        for (int t = t_begin; t < t_end; ++t) {
            const TileCoord c = decode_tile(t, L, P.tile_taps);
            if (L.stats.part && c.b != stat_b) {
                if (stat_b >= 0) flush_stats(stat_b);
                stat_b = c.b;
            }
            const int p = Wp + 1 + c.pt * TILE_M + m;
            const int yp = p / Wp, xp = p - yp * Wp;
            const bool valid = xp >= 1 && xp <= L.D && yp >= 1 && yp <= L.D;
            mbar_wait(smem_u32(tfull + ab), aphase);
            tc_fence_after();
#pragma unroll
            for (int ci = 0; ci < 6; ++ci) {
                const int c0 = ci * 16;
                if (c0 < P.NOUT && (ci & 1) == chalf) {
                    uint32_t r[16];
                    const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(ab * 2 * P.NOUT + c0);
                    tc_ld16(taddr, r);
                    float v[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]) + bias_s[c0 + i];
                    if (dual) {
                        tc_ld16(taddr + (uint32_t)P.NOUT, r);
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[i] += __uint_as_float(r[i]);
                    }
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = valid ? v[i] : 0.f;
                    if (P.epi == EPI_HEAD) {
                        if (valid) {
                            float *o = reinterpret_cast<float *>(L.out);
                            const size_t nv = (size_t)L.D * L.D * L.D;
                            const size_t vox = ((size_t)c.z * L.D + (yp - 1)) * L.D + (xp - 1);
#pragma unroll
                            for (int i = 0; i < 16; ++i)
                                if (c0 + i < L.Kout) o[((size_t)c.b * L.Kout + c0 + i) * nv + vox] = v[i];
                        }
                    } else {
                        uint32_t pk[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
                            pk[i] = *reinterpret_cast<uint32_t *>(&h2);
                        }
                        uint4 *o = reinterpret_cast<uint4 *>(L.out);
                        const int j = c0 >> 3;
                        if (P.epi == EPI_RAW) {
                            if (p < PP) {
                                const size_t base = (((size_t)c.b * L.CJ_out + j) * Wp + (c.z + 1)) * plane16 + p;
                                o[base] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                                o[base + (size_t)Wp * plane16] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                            }
                        } else if (valid) {                           // EPI_CONVT: scatter to output parity c.tap
                            const int D2 = 2 * L.D, Wp2 = D2 + 2;
                            const int a = c.tap >> 2, bq = (c.tap >> 1) & 1, cq = c.tap & 1;
                            const size_t plane2 = (size_t)Wp2 * Wp2;
                            const size_t pos2 = (size_t)(2 * (yp - 1) + bq + 1) * Wp2 + (2 * (xp - 1) + cq + 1);
                            const size_t base = (((size_t)c.b * L.CJ_out + j) * Wp2 + (2 * c.z + a + 1)) * plane2 + pos2;
                            o[base] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                            o[base + (size_t)Wp2 * plane2] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                        }
                        if (L.stats.part) {                            // column sums through a padded smem tile
#pragma unroll
                            for (int i = 0; i < 16; ++i) tile[lane * 17 + i] = v[i];
                            __syncwarp();
                            float s4[4] = {0.f, 0.f, 0.f, 0.f};              // four independent chains instead of one of 32
#pragma unroll
                            for (int rr = 0; rr < 32; ++rr) { const float x = tile[rr * 17 + col]; s4[rr & 3] += which ? x * x : x; }
                            acc_stat[ci] += (s4[0] + s4[1]) + (s4[2] + s4[3]);
                            __syncwarp();
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(tempty + ab));
            if (++ab == nbuf) { ab = 0; aphase ^= 1; }
        }
        if (L.stats.part && stat_b >= 0) flush_stats(stat_b);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == TC_PIPES) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------
// helper kernels of the bf16 path (all elementwise / HBM-bound)
// ------------------------------------------------------------------------------------------------
// pack PyTorch fp32 weights to [tap][KC][NOUT][8] bf16 (zero padded); ci runs over groups*8*KC_g
__global__ void tc_pack_weights_kernel(const float *__restrict__ src, __nv_bfloat16 *__restrict__ dst, int cout, int cin,
                                       int taps, int transposed, int KC, int NOUT)
{
    const int n = taps * KC * NOUT * 8;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const int i = e & 7; int r = e >> 3;
    const int co = r % NOUT; r /= NOUT;
    const int j = r % KC;
    const int t = r / KC;
    const int ci = j * 8 + i;
    float v = 0.f;
    if (co < cout && ci < cin) v = transposed ? src[((size_t)ci * cout + co) * taps + t] : src[((size_t)co * cin + ci) * taps + t];
    dst[e] = __float2bfloat16_rn(v);
}

__global__ void tc_pad_bias_kernel(const float *__restrict__ src, float *__restrict__ dst, int cout, int NOUT)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < NOUT) dst[i] = i < cout ? src[i] : 0.f;
}

// zero everything of a BP/PS tensor that no kernel writes: planes 0 and D+1, rows 0 and D+1, columns 0 and D+1
__global__ void __launch_bounds__(128)
tc_zero_border_kernel(uint4 *__restrict__ t, int D)
{
    const int Wp = D + 2, PP = Wp * Wp;
    const int zp = blockIdx.x % Wp;
    uint4 *plane = t + ((size_t)blockIdx.y * Wp + zp) * PP;
    const uint4 z4 = make_uint4(0, 0, 0, 0);
    if (zp == 0 || zp == Wp - 1) {
        for (int p = threadIdx.x; p < PP; p += blockDim.x) plane[p] = z4;
    } else {
        for (int p = threadIdx.x; p < Wp; p += blockDim.x) { plane[p] = z4; plane[(size_t)(Wp - 1) * Wp + p] = z4; }
        for (int y = threadIdx.x; y < Wp; y += blockDim.x) { plane[(size_t)y * Wp] = z4; plane[(size_t)y * Wp + Wp - 1] = z4; }
    }
}

// the same for up to 12 tensors in one launch (one block per z-plane of any of them)
struct ZeroJobs { int n; int first[13]; int D[12]; uint4 *p[12]; };
__global__ void __launch_bounds__(128)
tc_zero_border_multi_kernel(const __grid_constant__ ZeroJobs J)
{
    int k = 0;
    while (k + 1 < J.n && (int)blockIdx.x >= J.first[k + 1]) ++k;
    const int D = J.D[k], Wp = D + 2, PP = Wp * Wp;
    const int pl = blockIdx.x - J.first[k];                          // plane index over [chunks_total][Wp]
    const int zp = pl % Wp;
    uint4 *plane = J.p[k] + (size_t)pl * PP;
    const uint4 z4 = make_uint4(0, 0, 0, 0);
    if (zp == 0 || zp == Wp - 1) {
        for (int q = threadIdx.x; q < PP; q += blockDim.x) plane[q] = z4;
    } else {
        for (int q = threadIdx.x; q < Wp; q += blockDim.x) { plane[q] = z4; plane[(size_t)(Wp - 1) * Wp + q] = z4; }
        for (int y = threadIdx.x; y < Wp; y += blockDim.x) { plane[(size_t)y * Wp] = z4; plane[(size_t)y * Wp + Wp - 1] = z4; }
    }
}

int tc_zero_border_launch(void *tensor, int chunks_total, int D, cudaStream_t st)
{
    JHN_LAUNCH("tc_zero_border_kernel", st,
               tc_zero_border_kernel<<<dim3(D + 2, chunks_total), 128, 0, st>>>((uint4 *)tensor, D));
    return JHN_OK;
}

// InstanceNorm (from the fused sum / sum-of-squares) [+ residual] [ReLU] [+ skip], in place on a BP tensor;
// optionally also writes the parity-split copy the following stride-2 convolution reads.  HBM-bound: one block
// per (sample, 8-channel chunk, z-plane) walks the plane's contiguous run of padded positions with 16-byte
// accesses; the 8 (mean, rstd) pairs are computed once per block, pad columns are skipped (they stay zero).
#ifndef NORM_UNROLL_V
#define NORM_UNROLL_V 3
#endif
#ifndef NORM_MINB
#define NORM_MINB 1
#endif
#ifndef NORM_ZB_V
#define NORM_ZB_V 4
#endif
constexpr int NORM_THREADS = 256, NORM_UNROLL = NORM_UNROLL_V, NORM_ZB = NORM_ZB_V;      // NORM_ZB z-planes per block: the statistics prologue is paid once per 4 planes
template <bool HAS_RES, bool HAS_POST, bool HAS_PS>
__global__ void __launch_bounds__(NORM_THREADS, NORM_MINB)
tc_norm_act_kernel(uint4 *__restrict__ x, const StatPart stats, const uint4 *__restrict__ residual,
                   const uint4 *__restrict__ post_add, uint4 *__restrict__ ps_out, int relu, int D, int CJ, int NOUT,
                   float inv_count, float eps, uint32_t wp_magic)
{
    __shared__ float sc[16];                                          // mean[8], rstd[8]
    __shared__ double red[32][8][2];
    const int bj = blockIdx.y, b = bj / CJ, j = bj - b * CJ;
    const int Wp = D + 2, PP = Wp * Wp;
    {
        // second stage of the statistics: the producing kernel's per-CTA slots of sample b, added in slot order in fp64
        // (strided over 32 thread groups, then a fixed-order sum over the groups): identical bits on every run
        const int c0 = stat_owner((long long)b * stats.Tb, stats.grid, stats.T);
        const int c1 = stat_owner((long long)(b + 1) * stats.Tb - 1, stats.grid, stats.T);
        const int nslots = (c1 - c0 + 1) * 4;
        const int ch = threadIdx.x & 7, grp = threadIdx.x >> 3;
        const float2 *src = reinterpret_cast<const float2 *>(stats.part) + (size_t)(c0 + b) * 4 * NOUT + j * 8 + ch;
        double s1 = 0.0, s2 = 0.0;
        for (int s = grp; s < nslots; s += NORM_THREADS / 8) { const float2 v = src[(size_t)s * NOUT]; s1 += (double)v.x; s2 += (double)v.y; }
        red[grp][ch][0] = s1; red[grp][ch][1] = s2;
    }
    __syncthreads();
    if (threadIdx.x < 8) {
        double s1 = 0.0, s2 = 0.0;
        for (int g = 0; g < NORM_THREADS / 8; ++g) { s1 += red[g][threadIdx.x][0]; s2 += red[g][threadIdx.x][1]; }
        const double mean = s1 * (double)inv_count;
        const double var = fmax(s2 * (double)inv_count - mean * mean, 0.0);
        sc[threadIdx.x] = (float)mean;
        sc[8 + threadIdx.x] = (float)(1.0 / sqrt(var + (double)eps));
    }
    __syncthreads();
    float mean[8], rstd[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { mean[i] = sc[i]; rstd[i] = sc[8 + i]; }
    const int p_begin = Wp + 1, p_end = D * Wp + D + 1;               // first / one-past-last interior position
    const int Dh = D / 2, Wh = Dh + 2;
    const int z_end = min(D, ((int)blockIdx.x + 1) * NORM_ZB);
    for (int z = blockIdx.x * NORM_ZB; z < z_end; ++z) {
    const size_t plane = ((size_t)bj * Wp + z + 1) * PP;
    for (int p0 = p_begin + threadIdx.x; p0 < p_end; p0 += NORM_THREADS * NORM_UNROLL) {
        uint4 raw[NORM_UNROLL], rs[NORM_UNROLL], pa[NORM_UNROLL];
        bool ok[NORM_UNROLL];
        int yp[NORM_UNROLL], xp[NORM_UNROLL];
#pragma unroll
        for (int u = 0; u < NORM_UNROLL; ++u) {
            const int p = p0 + u * NORM_THREADS;
            yp[u] = (int)__umulhi((uint32_t)p, wp_magic);             // p / Wp (exact for p * Wp < 2^32)
            xp[u] = p - yp[u] * Wp;
            ok[u] = p < p_end && xp[u] >= 1 && xp[u] <= D;
            if (ok[u]) {
                raw[u] = x[plane + p];
                if (HAS_RES) rs[u] = residual[plane + p];
                if (HAS_POST) pa[u] = post_add[plane + p];
            }
        }
#pragma unroll
        for (int u = 0; u < NORM_UNROLL; ++u) {
            if (!ok[u]) continue;
            const int p = p0 + u * NORM_THREADS;
            const uint32_t w[4] = {raw[u].x, raw[u].y, raw[u].z, raw[u].w};
            const uint32_t rw[4] = {rs[u].x, rs[u].y, rs[u].z, rs[u].w};
            const uint32_t pw[4] = {pa[u].x, pa[u].y, pa[u].z, pa[u].w};
            uint32_t outw[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float f0 = (__uint_as_float(w[i] << 16) - mean[2 * i]) * rstd[2 * i];
                float f1 = (__uint_as_float(w[i] & 0xffff0000u) - mean[2 * i + 1]) * rstd[2 * i + 1];
                if (HAS_RES) { f0 += __uint_as_float(rw[i] << 16); f1 += __uint_as_float(rw[i] & 0xffff0000u); }
                if (relu) { f0 = fmaxf(f0, 0.f); f1 = fmaxf(f1, 0.f); }
                if (HAS_POST) { f0 += __uint_as_float(pw[i] << 16); f1 += __uint_as_float(pw[i] & 0xffff0000u); }
                __nv_bfloat162 h2 = __floats2bfloat162_rn(f0, f1);
                outw[i] = *reinterpret_cast<uint32_t *>(&h2);
            }
            const uint4 res = make_uint4(outw[0], outw[1], outw[2], outw[3]);
            x[plane + p] = res;
            if (HAS_PS) {                                              // [b][s][j][zp][pp] on the D/2 grid
                const int y = yp[u] - 1, xq = xp[u] - 1;
                const int s = ((z & 1) * 2 + (y & 1)) * 2 + (xq & 1);
                const size_t po = ((((size_t)b * 8 + s) * CJ + j) * Wh + (z >> 1) + 1) * Wh * Wh + (size_t)((y >> 1) + 1) * Wh + (xq >> 1) + 1;
                ps_out[po] = res;
            }
        }
    }
    }
}

// NCDHW fp32 [B][C][G^3] -> PS bf16 [b][s][j][zp][pp][8] on the G/2 grid (entry for jhn_v2v_forward with fp32 input)
__global__ void __launch_bounds__(256)
tc_ncdhw_to_ps_kernel(const float *__restrict__ in, uint4 *__restrict__ out, int C, int G, int CJ)
{
    const size_t nv = (size_t)G * G * G;
    const size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nv) return;
    const int bj = blockIdx.y, b = bj / CJ, j = bj - b * CJ;
    const int I = (int)(v / ((size_t)G * G)), r = (int)(v - (size_t)I * G * G), J = r / G, Kz = r - J * G;
    uint32_t pk[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c0 = j * 8 + 2 * i;
        const float a = c0 < C ? in[((size_t)b * C + c0) * nv + v] : 0.f;
        const float bb = c0 + 1 < C ? in[((size_t)b * C + c0 + 1) * nv + v] : 0.f;
        __nv_bfloat162 h2 = __floats2bfloat162_rn(a, bb);
        pk[i] = *reinterpret_cast<uint32_t *>(&h2);
    }
    const int Dh = G / 2, Wh = Dh + 2;
    const int s = ((I & 1) * 2 + (J & 1)) * 2 + (Kz & 1);
    const size_t po = ((((size_t)b * 8 + s) * CJ + j) * Wh + (I >> 1) + 1) * Wh * Wh + (size_t)((J >> 1) + 1) * Wh + (Kz >> 1) + 1;
    out[po] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
}

// BP bf16 -> NCDHW fp32 (debug / per-layer parity tests)
__global__ void __launch_bounds__(256)
tc_bp_to_ncdhw_kernel(const uint4 *__restrict__ in, float *__restrict__ out, int C, int D, int CJ)
{
    const int nv = D * D * D;
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nv) return;
    const int bj = blockIdx.y, b = bj / CJ, j = bj - b * CJ;
    const int z = v / (D * D), r = v - z * D * D, y = r / D, xq = r - y * D;
    const int Wp = D + 2;
    const uint4 raw = in[(((size_t)bj * Wp + z + 1) * Wp + y + 1) * Wp + xq + 1];
    const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c0 = j * 8 + 2 * i;
        if (c0 < C) out[((size_t)b * C + c0) * nv + v] = __uint_as_float(w[i] << 16);
        if (c0 + 1 < C) out[((size_t)b * C + c0 + 1) * nv + v] = __uint_as_float(w[i] & 0xffff0000u);
    }
}

// NCDHW fp32 -> BP bf16 interior (debug / per-layer parity tests)
__global__ void __launch_bounds__(256)
tc_ncdhw_to_bp_kernel(const float *__restrict__ in, uint4 *__restrict__ out, int C, int D, int CJ)
{
    const int nv = D * D * D;
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nv) return;
    const int bj = blockIdx.y, b = bj / CJ, j = bj - b * CJ;
    const int z = v / (D * D), r = v - z * D * D, y = r / D, xq = r - y * D;
    const int Wp = D + 2;
    uint32_t pk[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c0 = j * 8 + 2 * i;
        const float a = c0 < C ? in[((size_t)b * C + c0) * nv + v] : 0.f;
        const float bb = c0 + 1 < C ? in[((size_t)b * C + c0 + 1) * nv + v] : 0.f;
        __nv_bfloat162 h2 = __floats2bfloat162_rn(a, bb);
        pk[i] = *reinterpret_cast<uint32_t *>(&h2);
    }
    out[(((size_t)bj * Wp + z + 1) * Wp + y + 1) * Wp + xq + 1] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static inline int pad16(int c) { return (c + 15) / 16 * 16; }
static inline int tiles_per_plane(int D) { return cdiv((long long)(D - 1) * (D + 2) + D, TILE_M); }

// stacked-x-tap kernel for the C->C 3x3x3 layers (conv3_tc.cu)
size_t c3_weight_bytes(int NOUT);
bool c3_plan(int NOUT, int D, int max_smem, int *NS, int *PB);
int c3_pack(const float *src, __nv_bfloat16 *dst, int cout, int cin, int NOUT, cudaStream_t st);
int c3_launch(int NOUT, const void *in, const __nv_bfloat16 *w, const float *bias, void *out, StatPart *stats, int B, int D,
              int NS, int PB, int sms, int add_bias, cudaStream_t st);

struct TcLayer {
    TcProgram prog;
    __nv_bfloat16 *w;
    __nv_bfloat16 *w3;                                // [dz*3+dy][KC][dx*NOUT+co][8] for the stacked kernel, or null
    float *bias;
    int cin_pad, cout_pad, cout;
    size_t smem_bytes;
};

struct TcNet {
    TcLayer layer[NUM_LAYERS];
    void *blob;
    int max_smem;
    bool legacy_k3;                                   // JHN_CONV3_LEGACY=1: tap-per-MMA kernel for A/B measurements
};

static bool c3_eligible(const LayerDesc &d, int kind, int cin_pad, int cout_pad)
{
    return kind == 0 && cin_pad == cout_pad && cout_pad <= 80 && d.ks == 3;
}

// Build the stage program of one layer for output/position grid side D.
// kind: 0 = k3 s1 (BP in), 1 = k3 s2 (PS in), 2 = k2 s2 (PS in), 3 = convT k2 s2 (BP in), 4 = 1x1x1 (BP in)
static int build_program(TcProgram &P, int kind, int cin_pad, int cout_pad, int D, int max_smem)
{
    memset(&P, 0, sizeof(P));
    const int Wp = D + 2;
    P.KC = cin_pad / 8;
    P.NOUT = cout_pad;
    P.tile_taps = 1;
    P.epi = EPI_RAW;
    const int tap_bytes = P.KC * P.NOUT * 16;
    auto stage = [&](int chunk0, int dz, int pos_off) -> TcStage & {
        TcStage &S = P.st[P.nstages++];
        S.chunk0 = (int16_t)chunk0; S.dz = (int16_t)dz; S.pos_off = pos_off; S.ntaps = 0; S.wtap0 = 0;
        return S;
    };
    auto tap = [&](TcStage &S, int aoff, int widx) {
        if (S.ntaps == 0) S.wtap0 = (int16_t)widx;
        S.taps[S.ntaps].aoff = (int16_t)aoff; S.taps[S.ntaps].widx = (int16_t)widx; S.ntaps++;
    };
    if (kind == 0) {
        P.ntaps_total = 27;
        const bool resident = (size_t)27 * tap_bytes + 3 * ((size_t)P.KC * (TILE_M + 2 * Wp + 2) * 16) + TC_SMEM_TAIL <= (size_t)max_smem;
        P.resident = resident ? 1 : 0;
        if (resident) {                       // one box per z-plane, 9 in-plane taps each
            P.PB = TILE_M + 2 * Wp + 2;
            for (int dz = 0; dz < 3; ++dz) {
                TcStage &S = stage(0, dz, -(Wp + 1));
                for (int dy = 0; dy < 3; ++dy)
                    for (int dx = 0; dx < 3; ++dx) tap(S, (Wp + 1) + (dy - 1) * Wp + (dx - 1), (dz * 3 + dy) * 3 + dx);
            }
        } else {                              // weights streamed with the activations: one box per (dz,dy) row
            P.PB = TILE_M + 2;
            for (int dz = 0; dz < 3; ++dz)
                for (int dy = 0; dy < 3; ++dy) {
                    TcStage &S = stage(0, dz, (dy - 1) * Wp - 1);
                    for (int dx = 0; dx < 3; ++dx) tap(S, dx, (dz * 3 + dy) * 3 + dx);
                }
        }
    } else if (kind == 1) {                   // k3 s2 p1 on the parity-split input
        P.ntaps_total = 27; P.resident = 1; P.PB = TILE_M + Wp + 1;
        for (int s = 0; s < 8; ++s) {
            const int pz = s >> 2, py = (s >> 1) & 1, px = s & 1;
            for (int dz = (pz ? 0 : 1); dz <= 1; ++dz) {
                const int tz = pz ? (dz == 0 ? 0 : 2) : 1;
                TcStage &S = stage(s * P.KC, dz, -(Wp + 1));
                for (int ty = 0; ty < 3; ++ty) {
                    if ((ty == 1) != (py == 0)) continue;
                    for (int tx = 0; tx < 3; ++tx) {
                        if ((tx == 1) != (px == 0)) continue;
                        tap(S, (Wp + 1) + (ty == 0 ? -Wp : 0) + (tx == 0 ? -1 : 0), (tz * 3 + ty) * 3 + tx);
                    }
                }
            }
        }
    } else if (kind == 2) {                   // k2 s2 p0 on the parity-split input: 8 sub-volumes, one tap each
        P.ntaps_total = 8; P.resident = 1; P.PB = TILE_M;
        for (int s = 0; s < 8; ++s) { TcStage &S = stage(s * P.KC, 1, 0); tap(S, 0, s); }
    } else if (kind == 3) {                   // ConvTranspose k2 s2: tile index carries the output parity
        P.ntaps_total = 8; P.resident = 1; P.PB = TILE_M; P.tile_taps = 8; P.epi = EPI_CONVT;
        TcStage &S = stage(0, 1, 0); tap(S, 0, 0);
    } else {                                  // 1x1x1 head
        P.ntaps_total = 1; P.resident = 1; P.PB = TILE_M; P.epi = EPI_HEAD;
        TcStage &S = stage(0, 1, 0); tap(S, 0, 0);
    }
    int max_taps = 0;
    for (int s = 0; s < P.nstages; ++s) max_taps = max_taps > P.st[s].ntaps ? max_taps : P.st[s].ntaps;
    P.stage_bytes_a = (int)align_up((size_t)P.KC * P.PB * 16, 128);
    P.stage_bytes = P.stage_bytes_a + (P.resident ? 0 : (int)align_up((size_t)max_taps * tap_bytes, 128));
    P.w_bytes = P.resident ? (int)align_up((size_t)P.ntaps_total * tap_bytes, 128) : 0;
    const int fixed = P.w_bytes + TC_SMEM_TAIL;
    int slots = (max_smem - fixed) / P.stage_bytes;
    if (slots > TC_MAX_SLOTS) slots = TC_MAX_SLOTS;
    int n_mma = 0;
    for (int s = 0; s < P.nstages; ++s) n_mma += P.st[s].ntaps * (P.KC / 2);
    while (slots > 2 && slots * n_mma > MAX_DESC) --slots;
    if (slots * n_mma > MAX_DESC) return fail(JHN_ERR_SHAPE, "tensor-core conv: stage program too long (%d MMAs per tile)", n_mma);
    if (slots < 2) return fail(JHN_ERR_SHAPE, "tensor-core conv: tile does not fit shared memory (grid side %d)", D);
    P.nslots = slots;
    return JHN_OK;
}

static size_t program_smem(const TcProgram &P)
{
    return (size_t)P.w_bytes + (size_t)P.nslots * P.stage_bytes + TC_SMEM_TAIL;
}

static const int kLayerKind[NUM_LAYERS] = {1, 0, 0, 2, 0, 0, 3, 0, 0, 0, 0, 4};

int tc_create(jhn_v2v *net, const float *const *tensors, cudaStream_t st)
{
    int dev = 0, max_smem = 0;
    JHN_CUDA(cudaGetDevice(&dev));
    JHN_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    JHN_CUDA(cudaFuncSetAttribute(tc_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    TcNet *tc = new TcNet();
    // Leave a sliver of the SM's shared memory unclaimed: every resident CTA costs 1 KB of system-reserved shared memory, so a
    // convolution CTA sized to the full 227 KB keeps ANY other CTA off its SM — including the 0-byte transfer kernel of the
    // end-to-end path (jhn_pull_heatmap_boxes), which must run next to these persistent kernels to overlap PCIe with compute.
    static const int reserve = [] { const char *e = getenv("JHN_SMEM_RESERVE"); return e ? atoi(e) : 3072; }();
    tc->blob = nullptr; tc->max_smem = max_smem - reserve;
    const char *leg = getenv("JHN_CONV3_LEGACY");
    tc->legacy_k3 = leg && leg[0] == '1';
    net->tc = tc;
    size_t total = 0;
    for (int l = 0; l < NUM_LAYERS; ++l) {
        const LayerDesc &d = net->desc[l];
        TcLayer &T = tc->layer[l];
        T.cin_pad = pad16(d.cin); T.cout_pad = pad16(d.cout); T.cout = d.cout;
        const int taps = d.ks * d.ks * d.ks;
        total += align_up((size_t)taps * T.cin_pad * T.cout_pad * 2, 256) + align_up((size_t)T.cout_pad * 4, 256);
        if (c3_eligible(d, kLayerKind[l], T.cin_pad, T.cout_pad)) total += align_up(c3_weight_bytes(T.cout_pad), 256);
    }
    JHN_CUDA(cudaMalloc(&tc->blob, total));
    char *p = (char *)tc->blob;
    for (int l = 0; l < NUM_LAYERS; ++l) {
        const LayerDesc &d = net->desc[l];
        TcLayer &T = tc->layer[l];
        const int taps = d.ks * d.ks * d.ks;
        T.w = (__nv_bfloat16 *)p; p += align_up((size_t)taps * T.cin_pad * T.cout_pad * 2, 256);
        T.bias = (float *)p; p += align_up((size_t)T.cout_pad * 4, 256);
        T.w3 = nullptr;
        if (c3_eligible(d, kLayerKind[l], T.cin_pad, T.cout_pad)) {
            T.w3 = (__nv_bfloat16 *)p; p += align_up(c3_weight_bytes(T.cout_pad), 256);
            JHN_TRY(c3_pack(tensors[2 * l], T.w3, d.cout, d.cin, T.cout_pad, st));
        }
        const int n = taps * T.cin_pad * T.cout_pad;
        JHN_LAUNCH("tc_pack_weights_kernel", st,
                   tc_pack_weights_kernel<<<cdiv(n, 256), 256, 0, st>>>(tensors[2 * l], T.w, d.cout, d.cin, taps, d.transposed,
                                                                        T.cin_pad / 8, T.cout_pad));
        JHN_LAUNCH("tc_pad_bias_kernel", st,
                   tc_pad_bias_kernel<<<1, 128, 0, st>>>(tensors[2 * l + 1], T.bias, d.cout, T.cout_pad));
    }
    return JHN_OK;
}

void tc_destroy(jhn_v2v *net)
{
    if (!net->tc) return;
    if (net->tc->blob) cudaFree(net->tc->blob);
    delete net->tc;
    net->tc = nullptr;
}

// ---- buffers of one forward ------------------------------------------------------------------
struct TcBuffers {
    uint4 *vol_ps;                 // PS input volume        [B][8][C1/8][h+2][(h+2)^2]
    uint4 *A, *Bq, *C, *Dd, *E;    // h-grid BP tensors      [B][C2/8][h+2][(h+2)^2]
    uint4 *Xps;                    // PS copy of front1 out  [B][8][C2/8][q+2][(q+2)^2]
    uint4 *Pq, *Q, *R;             // q-grid BP tensors      [B][C4/8][q+2][(q+2)^2]
    float *stats;                  // [11][stat_slots][96][2] per-CTA partial sums (tc_ptx.cuh)
};

static size_t bp_units(int B, int cpad, int D) { return (size_t)B * (cpad / 8) * (D + 2) * (D + 2) * (D + 2); }

static void carve(Arena &a, const jhn_v2v *net, int B, int G, TcBuffers &t, bool with_vol)
{
    const TcNet *tc = net->tc;
    const int h = G / 2, q = G / 4;
    const int c1 = tc->layer[L_FRONT0].cin_pad, c2 = tc->layer[L_FRONT0].cout_pad, c4 = tc->layer[L_POOL].cout_pad;
    t.vol_ps = with_vol ? a.take<uint4>(8 * bp_units(B, c1, h)) : nullptr;
    t.A = a.take<uint4>(bp_units(B, c2, h)); t.Bq = a.take<uint4>(bp_units(B, c2, h)); t.C = a.take<uint4>(bp_units(B, c2, h));
    t.Dd = a.take<uint4>(bp_units(B, c2, h)); t.E = a.take<uint4>(bp_units(B, c2, h));
    t.Xps = a.take<uint4>(8 * bp_units(B, c2, q));
    t.Pq = a.take<uint4>(bp_units(B, c4, q)); t.Q = a.take<uint4>(bp_units(B, c4, q)); t.R = a.take<uint4>(bp_units(B, c4, q));
    t.stats = a.take<float>((size_t)11 * stat_slots(256, B) * 96 * 2);
}

size_t tc_volume_bytes(const jhn_v2v *net, int B, int G)
{
    return 8 * bp_units(B, net->tc->layer[L_FRONT0].cin_pad, G / 2) * 16;
}

size_t tc_workspace(const jhn_v2v *net, int B, int G)
{
    Arena a(nullptr, 0);
    TcBuffers t;
    carve(a, net, B, G, t, true);
    return a.off;
}

namespace {
struct TcCtx {
    const jhn_v2v *net; int B; cudaStream_t st; int sms;
    bool keep_bias;                                   // debug hook: raw conv output incl. bias; forward: an InstanceNorm follows

    // in: BP or PS tensor with `chunks_in` chunks per sample on grid side D (the GEMM-row grid)
    int conv(int l, const uint4 *in, int chunks_in, int D, void *out, int chunks_out, StatPart *stats) const
    {
        const TcLayer &T = net->tc->layer[l];
        int ns = 0, pb = 0;
        if (T.w3 && !net->tc->legacy_k3 && chunks_in == T.cin_pad / 8 && chunks_out == T.cout_pad / 8 &&
            c3_plan(T.cout_pad, D, net->tc->max_smem, &ns, &pb))
            return c3_launch(T.cout_pad, in, T.w3, T.bias, out, stats, B, D, ns, pb, sms, keep_bias ? 1 : 0, st);
        TcProgram P;
        JHN_TRY(build_program(P, kLayerKind[l], T.cin_pad, T.cout_pad, D, net->tc->max_smem));
        P.KC_load = (net->desc[l].cin + 7) / 8 < P.KC ? (net->desc[l].cin + 7) / 8 : P.KC;
        TcLaunch L;
        L.in = in; L.w = T.w; L.bias = T.bias; L.out = out; L.B = B; L.D = D; L.CJ_in = chunks_in; L.CJ_out = chunks_out;
        L.NT = tiles_per_plane(D); L.total_tiles = B * D * L.NT * P.tile_taps; L.Kout = T.cout;
        const int grid = L.total_tiles < sms ? L.total_tiles : sms;
        if (stats) { stats->grid = grid; stats->Tb = D * L.NT * P.tile_taps; stats->T = L.total_tiles; L.stats = *stats; }
        else L.stats = StatPart{nullptr, 0, 0, 0};
        const char *name = kLayerKind[l] == 0 ? (P.resident ? "tc_conv_k3_resident" : "tc_conv_k3_streamed")
                           : kLayerKind[l] == 1 ? "tc_conv_front_k3s2" : kLayerKind[l] == 2 ? "tc_conv_pool_k2s2"
                           : kLayerKind[l] == 3 ? "tc_conv_up_convT" : "tc_conv_head_1x1";
        JHN_LAUNCH(name, st, tc_conv_kernel<<<grid, TC_THREADS, program_smem(P), st>>>(P, L));
        return JHN_OK;
    }
    int norm(uint4 *x, const StatPart &stats, int l, int D, const uint4 *residual, bool relu, const uint4 *post_add, uint4 *ps) const
    {
        const TcLayer &T = net->tc->layer[l];
        const int nv = D * D * D, CJ = T.cout_pad / 8, Wp = D + 2;
        const uint32_t magic = (uint32_t)((0x100000000ull + (unsigned)Wp - 1) / (unsigned)Wp);   // ceil(2^32 / Wp)
        const dim3 grid(cdiv(D, NORM_ZB), B * CJ);
        const float inv = 1.f / (float)nv;
#define JHN_NORM(R, P, S)                                                                                              \
        JHN_LAUNCH("tc_norm_act_kernel", st,                                                                           \
                   (tc_norm_act_kernel<R, P, S><<<grid, NORM_THREADS, 0, st>>>(x, stats, residual, post_add, ps, relu ? 1 : 0, D, CJ, \
                                                                               T.cout_pad, inv, 1e-5f, magic)))
        if (residual && post_add && !ps) JHN_NORM(true, true, false);
        else if (residual && !post_add && ps) JHN_NORM(true, false, true);
        else if (residual && !post_add && !ps) JHN_NORM(true, false, false);
        else if (!residual && !post_add && !ps) JHN_NORM(false, false, false);
        else return fail(JHN_ERR_ARG, "norm: unsupported operand combination");
#undef JHN_NORM
        return JHN_OK;
    }
    int zero_border(uint4 *t, int chunks_total, int D) const { return tc_zero_border_launch(t, chunks_total, D, st); }
};
}  // namespace

bool head_supported(int K, int cin_pad, int cout_pad, int D);
int head_centroid_launch(const void *in, const __nv_bfloat16 *w, const float *bias, int B, int D, int K, int cin_pad, int cout_pad,
                         float spacing, float roi, const float *center3D, float *points, float *conf, int32_t *argmax, void *acc,
                         int sms, int max_smem, cudaStream_t st);

// `tail` non-null: the output layer runs fused with the centroid tail and `out` is not written (may be null).
// `carveB` >= B: the workspace is carved for carveB frame sets (a caller that walks a batch in sub-batches keeps one
// carving, so the cached zero borders stay valid for a shorter last pass); 0 = B.
// `borders` : 1 / 0 = the caller knows the zero borders of the tensors in `ws` are intact / must be rewritten, -1 = look
// `ws` up in the network's cache.
int tc_forward(const jhn_v2v *net, const void *volume_in, int in_layout, int B, int G, float *out, void *ws, size_t ws_bytes,
               cudaStream_t st, const TailArgs *tail, int carveB, int borders)
{
    if (carveB < B) carveB = B;
    const TcNet *tc = net->tc;
    if (!tc) return fail(JHN_ERR_ARG, "network was not created with JHN_BF16");
    const int h = G / 2, q = G / 4;
    if (h + 2 > 63) return fail(JHN_ERR_SHAPE, "bf16 path supports grid sides up to 122; got %d", G);
    Arena a(ws, ws_bytes);
    TcBuffers t;
    carve(a, net, carveB, G, t, in_layout != JHN_VOL_V2V_BF16);
    if (!a.ok()) return fail(JHN_ERR_WORKSPACE, "v2v bf16 workspace: need %zu bytes, got %zu", a.off, ws_bytes);
    int dev = 0, sms = 0;
    JHN_CUDA(cudaGetDevice(&dev));
    JHN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    TcCtx c{net, B, st, sms, false};
    const int j1 = tc->layer[L_FRONT0].cin_pad / 8, j2 = tc->layer[L_FRONT0].cout_pad / 8, j4 = tc->layer[L_POOL].cout_pad / 8;
    const uint4 *vol = (const uint4 *)volume_in;
    const bool borders_valid = borders >= 0 ? borders != 0 : net->borders_cached(ws, jhn_v2v::border_sig(1 + in_layout, carveB, G, 0, 0));
    if (in_layout != JHN_VOL_V2V_BF16) {
        if (!borders_valid) JHN_TRY(c.zero_border(t.vol_ps, carveB * 8 * j1, h));
        const size_t nv = (size_t)G * G * G;
        JHN_LAUNCH("tc_ncdhw_to_ps_kernel", st,
                   tc_ncdhw_to_ps_kernel<<<dim3(cdiv(nv, 256), B * j1), 256, 0, st>>>((const float *)volume_in, t.vol_ps, net->K, G, j1));
        vol = t.vol_ps;
    }
    if (!borders_valid) {
        ZeroJobs J;
        uint4 *ptrs[9] = {t.A, t.Bq, t.C, t.Dd, t.E, t.Pq, t.Q, t.R, t.Xps};
        const int CB = carveB;
        const int chunks[9] = {CB * j2, CB * j2, CB * j2, CB * j2, CB * j2, CB * j4, CB * j4, CB * j4, CB * 8 * j2};
        const int sides[9] = {h, h, h, h, h, q, q, q, q};
        J.n = 9;
        int total = 0;
        for (int i = 0; i < 9; ++i) { J.first[i] = total; J.D[i] = sides[i]; J.p[i] = ptrs[i]; total += chunks[i] * (sides[i] + 2); }
        J.first[9] = total;
        JHN_LAUNCH("tc_zero_border_kernel", st, tc_zero_border_multi_kernel<<<total, 128, 0, st>>>(J));
    }
    if (sms > 256) return fail(JHN_ERR_ARCH, "device has %d SMs; the statistics slots are sized for <= 256", sms);
    StatPart sp[11];
    for (int i = 0; i < 11; ++i) sp[i] = StatPart{t.stats + (size_t)i * stat_slots(256, carveB) * 96 * 2, 0, 0, 0};
    auto S = [&](int i) { return &sp[i]; };

    JHN_TRY(c.conv(L_FRONT0, vol, 8 * j1, h, t.A, j2, S(0)));                         // front_layers.0   v2vnet.py:90
    JHN_TRY(c.norm(t.A, *S(0), L_FRONT0, h, nullptr, true, nullptr, nullptr));
    JHN_TRY(c.conv(L_FRONT1A, t.A, j2, h, t.Bq, j2, S(1)));                           // front_layers.1 (Res3DBlock)
    JHN_TRY(c.norm(t.Bq, *S(1), L_FRONT1A, h, nullptr, true, nullptr, nullptr));
    JHN_TRY(c.conv(L_FRONT1B, t.Bq, j2, h, t.C, j2, S(2)));
    JHN_TRY(c.norm(t.C, *S(2), L_FRONT1B, h, t.A, true, nullptr, t.Xps));              // x = C (+ PS copy for the pool)
    JHN_TRY(c.conv(L_SKIPA, t.C, j2, h, t.A, j2, S(3)));                              // skip_res1        :76
    JHN_TRY(c.norm(t.A, *S(3), L_SKIPA, h, nullptr, true, nullptr, nullptr));
    JHN_TRY(c.conv(L_SKIPB, t.A, j2, h, t.Bq, j2, S(4)));
    JHN_TRY(c.norm(t.Bq, *S(4), L_SKIPB, h, t.C, true, nullptr, nullptr));             // s = Bq
    JHN_TRY(c.conv(L_POOL, t.Xps, 8 * j2, q, t.Pq, j4, S(5)));                        // encoder_pool1    :77
    JHN_TRY(c.norm(t.Pq, *S(5), L_POOL, q, nullptr, true, nullptr, nullptr));
    JHN_TRY(c.conv(L_MIDA, t.Pq, j4, q, t.Q, j4, S(6)));                              // mid_res          :78
    JHN_TRY(c.norm(t.Q, *S(6), L_MIDA, q, nullptr, true, nullptr, nullptr));
    JHN_TRY(c.conv(L_MIDB, t.Q, j4, q, t.R, j4, S(7)));
    JHN_TRY(c.norm(t.R, *S(7), L_MIDB, q, t.Pq, true, nullptr, nullptr));
    JHN_TRY(c.conv(L_UP, t.R, j4, q, t.A, j2, S(8)));                                 // decoder_upsample1 :79
    JHN_TRY(c.norm(t.A, *S(8), L_UP, h, nullptr, true, nullptr, nullptr));
    JHN_TRY(c.conv(L_DECA, t.A, j2, h, t.Dd, j2, S(9)));                              // decoder_res1     :80
    JHN_TRY(c.norm(t.Dd, *S(9), L_DECA, h, nullptr, true, nullptr, nullptr));
    JHN_TRY(c.conv(L_DECB, t.Dd, j2, h, t.E, j2, S(10)));
    JHN_TRY(c.norm(t.E, *S(10), L_DECB, h, t.A, true, t.Bq, nullptr));                 // relu(.. + x) + skip   :81
    if (tail) {                                                                        // output_layer + model.py:73-87
        const TcLayer &H = tc->layer[L_HEAD];
        return head_centroid_launch(t.E, H.w, H.bias, B, h, net->K, H.cin_pad, H.cout_pad, tail->spacing, tail->roi, tail->center3D,
                                    tail->points, tail->conf, tail->argmax, tail->acc, sms, tc->max_smem, st);
    }
    return c.conv(L_HEAD, t.E, j2, h, out, 0, nullptr);                               // output_layer     v2vnet.py:101
}

// Per-layer hook for the parity tests: fp32 NCDHW in -> (bf16 layout conversion) -> tensor-core conv ->
// raw (bias added, not normalised) output as fp32 NCDHW.  `D` is the layer's OUTPUT grid side, except for
// the transposed convolution where it is the INPUT grid side.
int tc_debug_layer(const jhn_v2v *net, int l, const float *in, int B, int D, float *out, void *ws, size_t ws_bytes, cudaStream_t st)
{
    const TcNet *tc = net->tc;
    if (!tc) return fail(JHN_ERR_ARG, "network was not created with JHN_BF16");
    if (l < 0 || l >= NUM_LAYERS) return fail(JHN_ERR_ARG, "layer %d out of range", l);
    const TcLayer &T = tc->layer[l];
    const LayerDesc &d = net->desc[l];
    const int kind = kLayerKind[l];
    const int Din = (kind == 1 || kind == 2) ? 2 * D : D;        // input grid side
    const int Dout = kind == 3 ? 2 * D : D;
    const int ji = T.cin_pad / 8, jo = T.cout_pad / 8;
    Arena a(ws, ws_bytes);
    const bool ps = kind == 1 || kind == 2;
    uint4 *tin = a.take<uint4>(ps ? 8 * bp_units(B, T.cin_pad, D) : bp_units(B, T.cin_pad, D));
    uint4 *tout = a.take<uint4>(bp_units(B, T.cout_pad, Dout));
    float *stats = a.take<float>((size_t)stat_slots(256, B) * 96 * 2);
    if (!a.ok()) return fail(JHN_ERR_WORKSPACE, "debug layer workspace: need %zu bytes, got %zu", a.off, ws_bytes);
    int dev = 0, sms = 0;
    JHN_CUDA(cudaGetDevice(&dev));
    JHN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    TcCtx c{net, B, st, sms, true};
    StatPart sp{stats, 0, 0, 0};
    JHN_TRY(c.zero_border(tin, B * (ps ? 8 : 1) * ji, D));
    if (ps) {
        const size_t nv = (size_t)Din * Din * Din;
        JHN_LAUNCH("tc_ncdhw_to_ps_kernel", st, tc_ncdhw_to_ps_kernel<<<dim3(cdiv(nv, 256), B * ji), 256, 0, st>>>(in, tin, d.cin, Din, ji));
    } else {
        const int nv = D * D * D;
        JHN_LAUNCH("tc_ncdhw_to_bp_kernel", st, tc_ncdhw_to_bp_kernel<<<dim3(cdiv(nv, 256), B * ji), 256, 0, st>>>(in, tin, d.cin, D, ji));
    }
    if (kind == 4) return c.conv(l, tin, ji, D, out, 0, nullptr);
    JHN_TRY(c.zero_border(tout, B * jo, Dout));
    JHN_TRY(c.conv(l, tin, (ps ? 8 : 1) * ji, D, tout, jo, &sp));
    const int nvo = Dout * Dout * Dout;
    JHN_LAUNCH("tc_bp_to_ncdhw_kernel", st, tc_bp_to_ncdhw_kernel<<<dim3(cdiv(nvo, 256), B * jo), 256, 0, st>>>(tout, out, d.cout, Dout, jo));
    return JHN_OK;
}

// Test hook: fp32 NCDHW activations -> BP bf16 -> fused output layer + centroid tail (head_tc.cu).
int tc_debug_head(const jhn_v2v *net, const float *in, int B, int h, const TailArgs &tail, void *ws, size_t ws_bytes, cudaStream_t st)
{
    const TcNet *tc = net->tc;
    const TcLayer &H = tc->layer[L_HEAD];
    if (!head_supported(net->K, H.cin_pad, H.cout_pad, h)) return fail(JHN_ERR_SHAPE, "fused head does not support K=%d", net->K);
    Arena a(ws, ws_bytes);
    const int ji = H.cin_pad / 8;
    uint4 *tin = a.take<uint4>(bp_units(B, H.cin_pad, h));
    char *acc = a.take<char>(head_acc_bytes(B, net->K));
    if (!a.ok()) return fail(JHN_ERR_WORKSPACE, "debug head workspace: need %zu bytes, got %zu", a.off, ws_bytes);
    int dev = 0, sms = 0;
    JHN_CUDA(cudaGetDevice(&dev));
    JHN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    JHN_TRY(tc_zero_border_launch(tin, B * ji, h, st));
    const int nv = h * h * h;
    JHN_LAUNCH("tc_ncdhw_to_bp_kernel", st,
               tc_ncdhw_to_bp_kernel<<<dim3(cdiv(nv, 256), B * ji), 256, 0, st>>>(in, tin, net->desc[L_HEAD].cin, h, ji));
    return head_centroid_launch(tin, H.w, H.bias, B, h, net->K, H.cin_pad, H.cout_pad, tail.spacing, tail.roi, tail.center3D,
                                tail.points, tail.conf, tail.argmax, acc, sms, tc->max_smem, st);
}

size_t tc_debug_workspace(const jhn_v2v *net, int l, int B, int D)
{
    const TcLayer &T = net->tc->layer[l];
    const int kind = kLayerKind[l];
    const int Dout = kind == 3 ? 2 * D : D;
    Arena a(nullptr, 0);
    const bool ps = kind == 1 || kind == 2;
    a.take<uint4>(ps ? 8 * bp_units(B, T.cin_pad, D) : bp_units(B, T.cin_pad, D));
    a.take<uint4>(bp_units(B, T.cout_pad, Dout));
    a.take<float>((size_t)stat_slots(256, B) * 96 * 2);
    return a.off;
}

}  // namespace jhn
