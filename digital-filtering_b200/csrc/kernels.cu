// csrc/kernels.cu -- the per-timestep hot path DIGITAL_FILTER::filter(dt) (df.cpp:449-468) as sm_100a kernels.
//
//   noise_kernel            H1  generate_white_noise   df.cpp:332-349  (replaced: counter-based pcg32, see noise.cuh)
//   ysweep_tma_kernel       H2y filtering_sweeps, y    df.cpp:360-383  tuned: TMA-staged slabs, dense band matrices
//   zsweep_epilogue_kernel  H2z filtering_sweeps, z    df.cpp:386-405  tuned: <ZK,1> recursive evaluation of the exponential
//                                                                      window (default), <ZK,0> direct Toeplitz register window
//                         + H3  correlate_fields       df.cpp:408-417
//                         + H4  apply_RST_scaling      df.cpp:419-447
//                         + H5  get_rho_T_fluc         df.cpp:470-485  (one epilogue, every output written once)
//   ysweep_simple_kernel / zsweep_epilogue_simple_kernel   one thread per cell, any per-cell half-width:
//                           the general (non row-uniform) path and the on-device cross-check of the tuned path
//   dfma_peak_kernel        fp64 roofline denominator
#include <cstdint>
#include <cstdlib>
#include <cstdio>
#include "noise.cuh"
#include "device.cuh"
#include "kernels.hpp"

// registers per thread of the recursive z-sweep (measured: capping it below 168 to fit a third CTA plus noise CTAs spills and loses)
#ifndef Z_REC_MAXNREG
#define Z_REC_MAXNREG 168
#endif
// pairs per thread of the noise kernel, NOISE_THREADS pairs apart (one position jump per thread, stride jumps between its pairs), two
// in flight at a time (NOISE_UNROLL: the pair transform is one long dependent chain; with the run-recursive y-sweep and the z-sweep
// resident beside it the noise kernel gets few warps per scheduler and needs the instruction-level parallelism).
// Measured ms/step on 1024x2048 profile / 4096x8192 profile: pairs 4 unroll 1: 0.1170 / 1.537; 4, 2: 0.1151 / 1.526; 4, 4: 0.1144 / 1.499;
// 8, 2: 0.1146 / 1.472
#ifndef NOISE_PAIRS
#define NOISE_PAIRS 8
#endif
#ifndef NOISE_UNROLL
#define NOISE_UNROLL 2
#endif
#ifndef NOISE_THREADS
#define NOISE_THREADS 128
#endif

namespace dfb {

// development aid (DFB_TIMELINE): earliest start / latest end of a launch on the GPU's global timer
__device__ __forceinline__ void tl_stamp(unsigned long long* tl, int end) {
    if (!tl) return;
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    if (end) atomicMax(tl + 1, gt); else atomicMin(tl, gt);
}

// =================================================================================================
// H1: white noise
// =================================================================================================
// One thread = NOISE_PAIRS pairs of one segment (row), NOISE_THREADS pairs apart, so that a warp's stores stay contiguous:
// ONE position jump per thread (segment jump o slot jump, two table reads), then from pair to pair a constant stride jump
// of 4*NOISE_THREADS draws -- one 64-bit multiply-add, like a plain LCG step -- instead of two table jumps per pair.
__device__ __forceinline__ void noise_thread(const NoiseParams& P, const PlaneDev& D, int slot0, int seg, int bz, int pl) {
    const NoiseArray& A = P.a[bz];
    if (seg >= A.n_seg) return;
    const uint64_t inc = P.pinc[pl][bz];
    // the table reads are requested together: one memory round trip per thread
    const int np = A.seg_np[seg];
    const Jump sj = reinterpret_cast<const Jump*>(A.seg_jump)[seg];
    const Jump tj = P.slot_jump[min(slot0, P.max_np - 1)];
    const int off = A.seg_off[seg];
    if (slot0 >= np) return;
    uint64_t s = sj.A * P.pstate[pl][bz] + inc * sj.C;
    s = tj.A * s + inc * tj.C;
    const uint64_t strideA = P.stride.A, strideC = inc * P.stride.C;
    const FieldDev& F = D.f[A.field];
    if (A.kind == 0) {
        // r_ys: pair `slot` of padded row `seg` holds the extended columns x0 = 2 slot + off and x0 + 1 (off = 2 q0 - row base)
        double* dst = F.r_ys + (size_t)pl * F.ps_ys + (size_t)seg * F.pitch_y;
        int x0 = 2 * slot0 + off;
        const bool vec_ok = (x0 & 1) == 0;                       // rows are 128-byte aligned: an even column is 16-byte aligned
        constexpr int NU = NOISE_UNROLL;
#pragma unroll NU
        for (int slot = slot0; slot < np; slot += NOISE_THREADS, x0 += 2 * NOISE_THREADS) {
            double z0, z1;
            normal_pair(s, inc, z0, z1);
            s = strideA * s + strideC;
            const bool in0 = x0 >= 0 && x0 < F.We, in1 = x0 + 1 >= 0 && x0 + 1 < F.We;
            if (in0 && in1 && vec_ok) {
                *reinterpret_cast<double2*>(dst + x0) = make_double2(z0, z1);
            } else {
                if (in0) dst[x0] = z0;
                if (in1) dst[x0 + 1] = z1;
            }
            if (slot + NOISE_THREADS >= slot0 + NOISE_PAIRS * NOISE_THREADS) break;
        }
    } else {
        // r_zs halo: element e = j*2M + h; h < M: global column h-M (left of the plane), else NzG + h-M
        const int M = F.Nz_max;
        double* dst = F.r_zs + (size_t)pl * F.ps_zs + (size_t)seg * F.pitch_z + F.zoff;
        const int Wz = D.W + 2 * M;
#pragma unroll 1
        for (int slot = slot0, r = 0; slot < np && r < NOISE_PAIRS; slot += NOISE_THREADS, ++r) {
            double z0, z1;
            normal_pair(s, inc, z0, z1);
            s = strideA * s + strideC;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int h = 2 * slot + i;
                if (h >= 2 * M) continue;
                const int g = h < M ? h - M : D.NzG + h - M;
                const int c = g - (D.k0 - M);
                if (c >= 0 && c < Wz) dst[c] = i ? z1 : z0;
            }
        }
    }
}

// one CTA per (NOISE_THREADS x NOISE_PAIRS pairs, segment, array x plane of the batch)
__global__ void __launch_bounds__(NOISE_THREADS) noise_kernel(const NoiseParams P, const PlaneDev D) {
    if (threadIdx.x == 0 && (blockIdx.y & 31) == 0) tl_stamp(P.tl, 0);     // sampled: one row in 32 (same-address atomics)
    noise_thread(P, D, (int)blockIdx.x * NOISE_PAIRS * NOISE_THREADS + (int)threadIdx.x, blockIdx.y, (int)(blockIdx.z % P.n_arrays), (int)(blockIdx.z / P.n_arrays));
    if (threadIdx.x == 0 && (blockIdx.y & 31) == 0) tl_stamp(P.tl, 1);
}

// =================================================================================================
// simple kernels: one thread per cell, coefficient row looked up by N.  Correct for any per-cell N.
// =================================================================================================
__global__ void __launch_bounds__(256) ysweep_simple_kernel(const PlaneDev D, int field) {
    const FieldDev& F = D.f[field];
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y, pl = blockIdx.z;
    if (x >= F.We) return;
    const int N = F.Ny_cell ? F.Ny_cell[(size_t)j * D.NzG + F.xk0 + x] : F.Ny_row[j];
    const double* b = D.coef_vals + D.coef_ptr[N] + N;
    const double* r = F.r_ys + (size_t)pl * F.ps_ys + (size_t)(j + F.Ny_max) * F.pitch_y + x;
    double sum = 0.0;
    for (int i = -N; i <= N; ++i) sum = fma(b[i], r[(ptrdiff_t)i * F.pitch_y], sum);   // df.cpp:373-375
    F.r_zs[(size_t)pl * F.ps_zs + (size_t)j * F.pitch_z + F.zoff + x + F.yshift] = sum;   // df.cpp:377
}

struct EpiOut { double u, v, w, T, rho; };

// H3 + H4 + H5 for one cell.  rc = the row's constants, z* = this step's z-sweep results.
__device__ __forceinline__ void epilogue_cell(const double* __restrict__ rc, const StepConsts& S,
                                              double zu, double zv, double zw, double fu, double fv, double fw,
                                              double& ou, double& ov, double& ow, EpiOut& o) {
    if (!S.first_step) {                                  // correlate_fields, df.cpp:415
        zu = fu * S.sa[0] + zu * S.sb[0];
        zv = fv * S.sa[1] + zv * S.sb[1];
        zw = fw * S.sa[2] + zw * S.sb[2];
    }
    ou = zu; ov = zv; ow = zw;                            // filt_old <- filt, df.cpp:440-442
    o.u = rc[0] * zu;                                     // df.cpp:436
    o.v = rc[1] * zu + rc[2] * zv;                        // df.cpp:437
    o.w = rc[3] * zw;                                     // df.cpp:438
    const double t2 = rc[4] * o.u;                        // df.cpp:478
    o.T = t2 * rc[5];                                     // df.cpp:480
    o.rho = -t2 * rc[6];                                  // df.cpp:481
}

__global__ void __launch_bounds__(256) zsweep_epilogue_simple_kernel(const PlaneDev D, const StepConsts S) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y, pl = blockIdx.z;
    if (k >= D.W) return;
    double z[3];
#pragma unroll
    for (int f = 0; f < 3; ++f) {
        const FieldDev& F = D.f[f];
        const int N = F.Nz_cell ? F.Nz_cell[(size_t)j * D.NzG + D.k0 + k] : F.Nz_row[j];
        const double* b = D.coef_vals + D.coef_ptr[N] + N;
        const double* r = F.r_zs + (size_t)pl * F.ps_zs + (size_t)j * F.pitch_z + F.zoff + F.Nz_max + k;
        double sum = 0.0;
        for (int i = -N; i <= N; ++i) sum = fma(b[i], r[i], sum);                        // df.cpp:397-399
        z[f] = sum;
    }
    const size_t idx = (size_t)pl * D.ps_cells + (size_t)j * D.W + k;
    EpiOut o;
    double ou, ov, ow;
    epilogue_cell(D.rowc + (size_t)j * ROWC, S, z[0], z[1], z[2],
                  D.f[0].filt_old[idx], D.f[1].filt_old[idx], D.f[2].filt_old[idx], ou, ov, ow, o);
    D.f[0].filt_old[idx] = ou; D.f[1].filt_old[idx] = ov; D.f[2].filt_old[idx] = ow;
    D.f[0].fluc[idx] = o.u; D.f[1].fluc[idx] = o.v; D.f[2].fluc[idx] = o.w;
    if (!S.first_step) { D.T_fluc[idx] = o.T; D.rho_fluc[idx] = o.rho; }                 // df.cpp:65 (quirk 3)
}

// =================================================================================================
// PTX helpers: mbarrier + TMA (cp.async.bulk[.tensor])
// =================================================================================================
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "W_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra D_%=;\n\t"
        "bra W_%=;\n\t"
        "D_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// the same, yielding the issue slot between polls: for waits that are expected to be long (a whole tile being staged) beside warps
// that have work to do
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    for (;;) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (done) break;
        __nanosleep(200);
    }
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// =================================================================================================
// H2y tuned: y-sweep for row-uniform half-widths.
//
// One CTA = Y_G = 4 row groups (32 consecutive output rows of one field) x 128 columns.
// Warp 4 is the TMA producer: it streams the union of the groups' input-row windows, chunk by
// chunk (RC = 8 padded rows x 128 columns, one cp.async.bulk.tensor.2d box), through a ring of NS
// shared-memory stages, together with the matching RC x 8 slice of each active group's dense band
// matrix (cp.async.bulk).  Sharing one sample stream between 4 groups cuts the L2 -> SM traffic
// from (8+2N)/8 to (32+2N)/32 loads per output.  Warps 0-3 are the consumers, one row group each:
// lane l owns columns {2l, 2l+1, 64+2l, 65+2l} and the group's 8 rows = 32 fp64 accumulators in
// registers; per input row it issues 2 LDS.128 (samples, conflict-free) + 4 LDS.128 (coefficients,
// warp-broadcast) for 32 DFMA.  out[jj] += C[t][jj] * x[t]: the band matrix holds b_{N(jj)}[t-..]
// and exact zeros outside each row's own half-width, so rows of different N share one pass
// (adding +0*x leaves a sum unchanged; noise is finite).
// =================================================================================================
template <int RC, int NS, int TK = Y_TK>
struct YSmem {
    double samples[NS][RC][TK];
    double coefs[NS][Y_G][RC][YJ];
    uint64_t full[NS];
    uint64_t empty[NS];
};

// TK = 128 columns per tile (lane owns 4) or 64 (lane owns 2: twice as many, half as heavy CTAs for planes that do not fill the GPU)
template <int RC, int NS, int TK>
__global__ void __launch_bounds__(160, 3) ysweep_tma_kernel(const __grid_constant__ YMaps maps, const YParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    YSmem<RC, NS, TK>& sm = *reinterpret_cast<YSmem<RC, NS, TK>*>(smem_raw);
    constexpr int NC = TK / 32;            // columns per lane
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // One tile per CTA (gridDim.x == n_tiles), or a resident grid whose CTAs walk the (longest-first) tile list with a stride:
    // with every CTA resident from the start, the block scheduler has nothing of this kernel pending and lets the next step's
    // noise CTAs (low-priority stream) in beside it.
    const int nP = P.D.P;                 // planes of a batch: tile list shared, plane = fastest index (longest-first order kept)
    for (int tix = blockIdx.x; tix < P.n_tiles * nP; tix += gridDim.x) {
    const YTile t = P.tiles[P.tile0 + tix / nP];
    const int pl = tix % nP;
    const FieldDev& F = P.D.f[t.field];
    const int prow0 = pl * F.rows_y;      // first padded row of this plane in the stacked r_ys

    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], Y_G); }
        mbar_fence_init();
        if (tix == 0 && P.zcounter) *P.zcounter = P.zcounter_init;      // the z-sweep that follows pulls its items from here
        tl_stamp(P.tl, 0);
    }
    __syncthreads();
    // Programmatic dependent launch: everything above overlapped the previous kernel's tail; from here on we touch
    // global data it may have produced (noise) or still read.
    if (tix == (int)blockIdx.x) asm volatile("griddepcontrol.wait;" ::: "memory");

    if (warp == Y_G) {
        if (lane == 0) {
            const CUtensorMap* map = &maps.m[t.field];
            int cs[Y_G], ce[Y_G];
            const double* cm[Y_G];
#pragma unroll
            for (int w = 0; w < Y_G; ++w) {
                if (w < t.ngroups) {
                    const YGroup g = P.groups[t.g0 + w];
                    cs[w] = g.cstart; ce[w] = g.cstart + g.nchunks; cm[w] = P.cmat + g.cmat_off;
                } else { cs[w] = 0; ce[w] = 0; cm[w] = nullptr; }
            }
            int i = 0;
            for (int c = t.cbegin; c < t.cend; ++c, ++i) {
                const int s = i % NS;
                if (i >= NS) mbar_wait(&sm.empty[s], ((i / NS) - 1) & 1);
                uint32_t bytes = (uint32_t)(sizeof(double) * RC * TK);
#pragma unroll
                for (int w = 0; w < Y_G; ++w) if (c >= cs[w] && c < ce[w]) bytes += (uint32_t)(sizeof(double) * RC * YJ);
                mbar_expect_tx(&sm.full[s], bytes);
                tma_load_2d(&sm.samples[s][0][0], map, t.col0, prow0 + c * RC, &sm.full[s]);
#pragma unroll
                for (int w = 0; w < Y_G; ++w)
                    if (c >= cs[w] && c < ce[w])
                        tma_load_1d(&sm.coefs[s][w][0][0], cm[w] + (long long)(c - cs[w]) * RC * YJ, RC * YJ * sizeof(double), &sm.full[s]);
            }
        }
    } else {

    const bool have = warp < t.ngroups;
    YGroup g{};
    if (have) g = P.groups[t.g0 + warp];
    const int my_cs = have ? g.cstart : 0, my_ce = have ? g.cstart + g.nchunks : 0;

    double acc[YJ][NC];
#pragma unroll
    for (int jj = 0; jj < YJ; ++jj)
#pragma unroll
        for (int q = 0; q < NC; ++q) acc[jj][q] = 0.0;

    long long tw = 0, tc = 0;
    const long long tstart = P.debug ? clock64() : 0;
    int i = 0;
    for (int c = t.cbegin; c < t.cend; ++c, ++i) {
        const int s = i % NS;
        const long long ta = P.debug ? clock64() : 0;
        mbar_wait(&sm.full[s], (i / NS) & 1);
        const long long tb = P.debug ? clock64() : 0;
        tw += tb - ta;
        if (c >= my_cs && c < my_ce) {
#pragma unroll
            for (int r = 0; r < RC; ++r) {
                const double2 xa = *reinterpret_cast<const double2*>(&sm.samples[s][r][2 * lane]);
                if (NC == 4) {
                    const double2 xb = *reinterpret_cast<const double2*>(&sm.samples[s][r][(TK / 2 + 2 * lane) % TK]);
#pragma unroll
                    for (int jj = 0; jj < YJ; jj += 2) {
                        const double2 cc = *reinterpret_cast<const double2*>(&sm.coefs[s][warp][r][jj]);
                        acc[jj][0] = fma(cc.x, xa.x, acc[jj][0]);
                        acc[jj][1] = fma(cc.x, xa.y, acc[jj][1]);
                        acc[jj][NC - 2] = fma(cc.x, xb.x, acc[jj][NC - 2]);
                        acc[jj][NC - 1] = fma(cc.x, xb.y, acc[jj][NC - 1]);
                        acc[jj + 1][0] = fma(cc.y, xa.x, acc[jj + 1][0]);
                        acc[jj + 1][1] = fma(cc.y, xa.y, acc[jj + 1][1]);
                        acc[jj + 1][NC - 2] = fma(cc.y, xb.x, acc[jj + 1][NC - 2]);
                        acc[jj + 1][NC - 1] = fma(cc.y, xb.y, acc[jj + 1][NC - 1]);
                    }
                } else {
#pragma unroll
                    for (int jj = 0; jj < YJ; jj += 2) {
                        const double2 cc = *reinterpret_cast<const double2*>(&sm.coefs[s][warp][r][jj]);
                        acc[jj][0] = fma(cc.x, xa.x, acc[jj][0]);
                        acc[jj][1] = fma(cc.x, xa.y, acc[jj][1]);
                        acc[jj + 1][0] = fma(cc.y, xa.x, acc[jj + 1][0]);
                        acc[jj + 1][1] = fma(cc.y, xa.y, acc[jj + 1][1]);
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.empty[s]);
        if (P.debug) tc += clock64() - tb;
    }
    const long long tloop = P.debug ? clock64() : 0;
    if (have) {

    // r_zs interior (df.cpp:377): extended column x -> logical column x + yshift
    const int xa0 = t.col0 + 2 * lane;
    const bool vec_ok = ((F.zoff + F.yshift) & 1) == 0;   // pitch_z is even, x is even
#pragma unroll
    for (int jj = 0; jj < YJ; ++jj) {
        if (jj >= g.nrows) break;
        double* dst = F.r_zs + (size_t)pl * F.ps_zs + (size_t)(g.j0 + jj) * F.pitch_z + F.zoff + F.yshift;
#pragma unroll
        for (int h = 0; h < NC / 2; ++h) {
            const int x = xa0 + 64 * h;
            if (vec_ok && x + 1 < F.We) {
                *reinterpret_cast<double2*>(dst + x) = make_double2(acc[jj][2 * h], acc[jj][2 * h + 1]);
            } else {
                if (x < F.We) dst[x] = acc[jj][2 * h];
                if (x + 1 < F.We) dst[x + 1] = acc[jj][2 * h + 1];
            }
        }
    }
    if (lane == 0) tl_stamp(P.tl, 1);
    if (P.debug && lane == 0) {
        const long long tend = clock64();
        atomicAdd(P.prof + 0, (unsigned long long)tw);                 // consumer waits on full barriers
        atomicAdd(P.prof + 1, (unsigned long long)tc);                 // chunk compute (incl. skipped chunks)
        atomicAdd(P.prof + 2, (unsigned long long)(tend - tloop));     // stores
        atomicAdd(P.prof + 3, (unsigned long long)(tend - tstart));    // whole tile (after setup)
        atomicAdd(P.prof + 4, 1ull);
        atomicAdd(P.prof + 5, (unsigned long long)(t.cend - t.cbegin));
    }
    }   // have
    }   // consumer
    __syncthreads();        // every warp is done with the ring and its barriers before the next tile re-initialises them
    }   // tiles
}

// =================================================================================================
// H2y tuned, recursive form: row groups whose rows all share one half-width N >= 16.
//
// The reference's coefficients are a truncated two-sided exponential, b_i = a^|i| / s with a = exp(-2 pi / N)
// (df.cpp:168-177).  Most rows of a group's window lie in EVERY output row's window and on one side of every output row:
// with i = position in the window (padded row w0 is i = 0, output row t is centred on i = N + t)
//   low  bulk, nrows-1 <= i <= N-1:   its weight for output t is a^(N+t-i) = a^(N+t-i_ll) * a^(i_ll-i)
//   high bulk, N+nrows <= i <= 2N:    its weight for output t is a^(i-N-t) = a^(i_fh-N-t) * a^(i-i_fh)
// so whole 8-row chunks of bulk rows are folded in a binary tree (a, a^2, a^4) into ONE running sum per column,
// G <- a^8 G + T or K <- K + p T, p <- a^8 p (8 FMAs per column and chunk instead of 64), and at the end
// out_t = dense_t + gl_t G + gh_t K with 16 host-computed factors per group, all <= 1 (no cancellation).  The chunks at the two ends
// of the window and around the output rows go through the dense band-matrix path exactly as in ysweep_tma_kernel.
// Same tiles, same TMA ring, same barriers; band-matrix slices are only fetched for the dense chunks.
// Measured against the oracle: <= 1e-14 of the rms (gate 1e-12).  A column's result does not depend on the slab it is computed in.
// =================================================================================================
template <int RC, int NS>
__global__ void __launch_bounds__(160, 3) ysweep_rec_kernel(const __grid_constant__ YMaps maps, const YParams P) {
    static_assert(RC == 8, "the chunk tree folds 8 rows");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    YSmem<RC, NS>& sm = *reinterpret_cast<YSmem<RC, NS>*>(smem_raw);
    const int nP = P.D.P;
    const YTile t = P.tiles[P.tile0 + blockIdx.x / nP];
    const int pl = blockIdx.x % nP;
    const FieldDev& F = P.D.f[t.field];
    const int prow0 = pl * F.rows_y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], Y_G); }
        mbar_fence_init();
        if (blockIdx.x == 0 && P.zcounter) *P.zcounter = P.zcounter_init;
        tl_stamp(P.tl, 0);
    }
    __syncthreads();
    asm volatile("griddepcontrol.wait;" ::: "memory");

    // chunk c of a group: 1 = low bulk, 2 = high bulk, 0 = dense (band matrix)
    auto chunk_kind = [](int rec, int c, int w0, int Nn, int nrows) {
        if (!rec) return 0;                 // mixed group: dense throughout
        const int i0 = c * RC - w0;
        if (i0 >= nrows - 1 && i0 + RC - 1 <= Nn - 1) return 1;
        if (i0 >= Nn + nrows && i0 + RC - 1 <= 2 * Nn) return 2;
        return 0;
    };

    if (warp == Y_G) {                     // producer
        if (lane == 0) {
            const CUtensorMap* map = &maps.m[t.field];
            int cs[Y_G], ce[Y_G], gw0[Y_G], gN[Y_G], gnr[Y_G], grec[Y_G];
            const double* cm[Y_G];
#pragma unroll
            for (int w = 0; w < Y_G; ++w) {
                if (w < t.ngroups) {
                    const YGroup g = P.groups[t.g0 + w];
                    cs[w] = g.cstart; ce[w] = g.cstart + g.nchunks; cm[w] = P.cmat + g.cmat_off; gw0[w] = g.w0; gN[w] = g.Nmax; gnr[w] = g.nrows; grec[w] = g.rec;
                } else { cs[w] = 0; ce[w] = 0; cm[w] = nullptr; gw0[w] = 0; gN[w] = 0; gnr[w] = 0; grec[w] = 0; }
            }
            int i = 0;
            for (int c = t.cbegin; c < t.cend; ++c, ++i) {
                const int s = i % NS;
                if (i >= NS) mbar_wait(&sm.empty[s], ((i / NS) - 1) & 1);
                bool need[Y_G];
                uint32_t bytes = (uint32_t)(sizeof(double) * RC * Y_TK);
#pragma unroll
                for (int w = 0; w < Y_G; ++w) {
                    need[w] = c >= cs[w] && c < ce[w] && chunk_kind(grec[w], c, gw0[w], gN[w], gnr[w]) == 0;
                    if (need[w]) bytes += (uint32_t)(sizeof(double) * RC * YJ);
                }
                mbar_expect_tx(&sm.full[s], bytes);
                tma_load_2d(&sm.samples[s][0][0], map, t.col0, prow0 + c * RC, &sm.full[s]);
#pragma unroll
                for (int w = 0; w < Y_G; ++w)
                    if (need[w])
                        tma_load_1d(&sm.coefs[s][w][0][0], cm[w] + (long long)(c - cs[w]) * RC * YJ, RC * YJ * sizeof(double), &sm.full[s]);
            }
        }
        return;
    }

    const bool have = warp < t.ngroups;
    YGroup g{};
    if (have) g = P.groups[t.g0 + warp];
    const int my_cs = have ? g.cstart : 0, my_ce = have ? g.cstart + g.nchunks : 0;
    const int Nn = g.Nmax, nrows = g.nrows;
    const double* par = P.yrec + (size_t)Nn * 16;
    double ra = 0.0, a2 = 0.0, a4 = 0.0, a8 = 0.0;
    if (have && g.rec) { ra = __ldg(par); a2 = __ldg(par + 10); a4 = __ldg(par + 11); a8 = __ldg(par + 12); }

    double acc[YJ][4];
#pragma unroll
    for (int jj = 0; jj < YJ; ++jj) { acc[jj][0] = acc[jj][1] = acc[jj][2] = acc[jj][3] = 0.0; }
    double G[4] = {0.0, 0.0, 0.0, 0.0}, K[4] = {0.0, 0.0, 0.0, 0.0}, pw = 1.0;

    int it = 0;
    for (int c = t.cbegin; c < t.cend; ++c, ++it) {
        const int s = it % NS;
        mbar_wait(&sm.full[s], (it / NS) & 1);
        if (c >= my_cs && c < my_ce) {
            const int kind = chunk_kind(g.rec, c, g.w0, Nn, nrows);
            if (kind == 1) {
                // ---- low bulk: G <- a^8 G + sum_r a^(7-r) x_r ----
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    double2 x[RC];
#pragma unroll
                    for (int r = 0; r < RC; ++r) x[r] = *reinterpret_cast<const double2*>(&sm.samples[s][r][64 * h + 2 * lane]);
#pragma unroll
                    for (int q = 0; q < RC / 2; ++q) { x[q].x = __fma_rn(ra, x[2 * q].x, x[2 * q + 1].x); x[q].y = __fma_rn(ra, x[2 * q].y, x[2 * q + 1].y); }
#pragma unroll
                    for (int q = 0; q < RC / 4; ++q) { x[q].x = __fma_rn(a2, x[2 * q].x, x[2 * q + 1].x); x[q].y = __fma_rn(a2, x[2 * q].y, x[2 * q + 1].y); }
                    x[0].x = __fma_rn(a4, x[0].x, x[1].x); x[0].y = __fma_rn(a4, x[0].y, x[1].y);
                    G[2 * h] = __fma_rn(a8, G[2 * h], x[0].x);
                    G[2 * h + 1] = __fma_rn(a8, G[2 * h + 1], x[0].y);
                }
            } else if (kind == 2) {
                // ---- high bulk: K += p * sum_r a^r x_r, p <- a^8 p ----
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    double2 x[RC];
#pragma unroll
                    for (int r = 0; r < RC; ++r) x[r] = *reinterpret_cast<const double2*>(&sm.samples[s][r][64 * h + 2 * lane]);
#pragma unroll
                    for (int q = 0; q < RC / 2; ++q) { x[q].x = __fma_rn(ra, x[2 * q + 1].x, x[2 * q].x); x[q].y = __fma_rn(ra, x[2 * q + 1].y, x[2 * q].y); }
#pragma unroll
                    for (int q = 0; q < RC / 4; ++q) { x[q].x = __fma_rn(a2, x[2 * q + 1].x, x[2 * q].x); x[q].y = __fma_rn(a2, x[2 * q + 1].y, x[2 * q].y); }
                    x[0].x = __fma_rn(a4, x[1].x, x[0].x); x[0].y = __fma_rn(a4, x[1].y, x[0].y);
                    K[2 * h] = __fma_rn(pw, x[0].x, K[2 * h]);
                    K[2 * h + 1] = __fma_rn(pw, x[0].y, K[2 * h + 1]);
                }
                pw = __dmul_rn(pw, a8);
            } else {
                // ---- window ends and the chunks around the output rows: dense band matrix, as in ysweep_tma_kernel ----
#pragma unroll
                for (int r = 0; r < RC; ++r) {
                    const double2 xa = *reinterpret_cast<const double2*>(&sm.samples[s][r][2 * lane]);
                    const double2 xb = *reinterpret_cast<const double2*>(&sm.samples[s][r][64 + 2 * lane]);
#pragma unroll
                    for (int jj = 0; jj < YJ; jj += 2) {
                        const double2 cc = *reinterpret_cast<const double2*>(&sm.coefs[s][warp][r][jj]);
                        acc[jj][0] = __fma_rn(cc.x, xa.x, acc[jj][0]);
                        acc[jj][1] = __fma_rn(cc.x, xa.y, acc[jj][1]);
                        acc[jj][2] = __fma_rn(cc.x, xb.x, acc[jj][2]);
                        acc[jj][3] = __fma_rn(cc.x, xb.y, acc[jj][3]);
                        acc[jj + 1][0] = __fma_rn(cc.y, xa.x, acc[jj + 1][0]);
                        acc[jj + 1][1] = __fma_rn(cc.y, xa.y, acc[jj + 1][1]);
                        acc[jj + 1][2] = __fma_rn(cc.y, xb.x, acc[jj + 1][2]);
                        acc[jj + 1][3] = __fma_rn(cc.y, xb.y, acc[jj + 1][3]);
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.empty[s]);
    }
    if (!have) return;

    // out_t = dense_t + gl_t G + gh_t K; then r_zs interior (df.cpp:377): extended column x -> logical column x + yshift
    const double* gc = P.ygc + (size_t)(g.rec ? g.gc_off : 0) * 16;      // (mixed groups: G = K = 0, the factors do not matter)
    const int xa0 = t.col0 + 2 * lane;
    const bool vec_ok = ((F.zoff + F.yshift) & 1) == 0;
#pragma unroll
    for (int jj = 0; jj < YJ; ++jj) {
        if (jj >= g.nrows) break;
        const double gl = __ldg(gc + jj), gh = __ldg(gc + 8 + jj);
        double* dst = F.r_zs + (size_t)pl * F.ps_zs + (size_t)(g.j0 + jj) * F.pitch_z + F.zoff + F.yshift;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int x = xa0 + 64 * h;
            const double v0 = __fma_rn(gh, K[2 * h], __fma_rn(gl, G[2 * h], acc[jj][2 * h]));
            const double v1 = __fma_rn(gh, K[2 * h + 1], __fma_rn(gl, G[2 * h + 1], acc[jj][2 * h + 1]));
            if (vec_ok && x + 1 < F.We) {
                *reinterpret_cast<double2*>(dst + x) = make_double2(v0, v1);
            } else {
                if (x < F.We) dst[x] = v0;
                if (x + 1 < F.We) dst[x + 1] = v1;
            }
        }
    }
    if (lane == 0) tl_stamp(P.tl, 1);
}

// =================================================================================================
// H2y tuned, run-recursive form: EVERY row group is evaluated through the exponential structure of the coefficients.
//
// b_i = a^|i| / s with a = exp(-2 pi / N) (df.cpp:168-177), so for a row j of half-width N
//     out_j = (F_j + B_j - x_j) / s,   F_j = sum_{i=0..N} a^i x_{j-i},   B_j = sum_{i=0..N} a^i x_{j+i},
// and inside a run of rows that share N both sums slide:  F_{j+1} = a F_j + x_{j+1} - a^(N+1) x_{j-N},
// B_{j-1} = a B_j + x_{j-1} - a^(N+1) x_{j+N}.  The half-width changes every few rows on boundary-layer grids, so the
// rows are cut into groups of consecutive rows of ONE half-width (a lone row is a group of one): up to 8 rows, F walking
// up from the first row and B down from the last, or -- N >= 25, runs longer than 8 -- up to min(32, 0.33 N) rows with
// both sums walking in lock step from the first row.  Per group and column: two Horner starts over N+1 samples each
// (run as four interleaved chains in a^4 per side: eight independent FMA chains per thread) + 4 FMAs per further row
// -- 2(N+1) + 4R - 2 FMAs where the direct sum spends R(2N+1); 5x fewer on the 1024x2048 boundary-layer profile, and no
// zero-padded band matrices at all.
//
// One tile = a block of up to 128 output rows x 32 columns of one field: its whole input window (block + 2 N_max rows)
// is staged ONCE in shared memory by TMA (32-row boxes of cp.async.bulk.tensor.2d), 256 bytes per row: a warp reads one
// row per LDS.64, conflict-free.  Persistent grid, one CTA per SM: a producer warp stages the next tile (window + the
// tile's group descriptors, one mbarrier) into the second buffer while 16 consumer warps pull the current tile's groups
// from a shared-memory counter (most expensive first); the consumers issue no global load at all.  Lanes are columns,
// so a column's arithmetic does not depend on the tile or slab it sits in.  With every CTA resident from the start the
// block scheduler has nothing of this kernel pending and lets the next step's noise CTAs in beside it.
// Bound by shared-memory bandwidth (one 8-byte sample per FMA), not by the fp64 pipe.
// Agreement with the reference's sequential sum: ~1e-15 of the rms (gate 1e-12).
// =================================================================================================
struct YRSmemCtl {                       // control block in front of the two window buffers
    uint64_t full[2], empty[2];
    int next_group[2];
    YRTile tile[2];
    int plane[2];
    int pad[2];
    alignas(16) YRGroup groups[2][YR_MAXG];      // cp.async.bulk destination
};

__global__ void __launch_bounds__(32 * (YR_CONSUMERS + 1), 1) ysweep_run_kernel(const __grid_constant__ YRMaps maps, const YParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    YRSmemCtl& ctl = *reinterpret_cast<YRSmemCtl*>(smem_raw);
    constexpr size_t CTL = (sizeof(YRSmemCtl) + 127) / 128 * 128;
    const size_t wbytes = (size_t)P.r_wrows * YR_C * sizeof(double);   // one window buffer (whole boxes)
    const int nbuf = P.r_nbuf;                                         // 2: the next tile is staged while this one is computed; 1: windows too tall for two
    const int nP = P.D.P;
    const int n_all = P.n_rtiles * nP;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < 2; ++s) { mbar_init(&ctl.full[s], 1); mbar_init(&ctl.empty[s], YR_CONSUMERS); }
        mbar_fence_init();
        if (blockIdx.x == 0 && P.zcounter) *P.zcounter = P.zcounter_init;      // the z-sweep that follows pulls its items from here
        tl_stamp(P.tl, 0);
    }
    __syncthreads();
    // Programmatic dependent launch: everything above overlapped the previous kernel's tail; the noise is read from here on
    asm volatile("griddepcontrol.wait;" ::: "memory");

    // Persistent, one CTA per SM.  The tiles (most expensive first) are handed out through a global counter: CTA b starts on tile b,
    // its producer claims every further tile while the consumers work (a static stride-gridDim.x walk gives CTA 0 the dearest
    // tile of every round: 10 % over the mean load on the 1024x2048 profile, 70 % on 512x512).  Only the producer knows the
    // sequence; the consumers take each tile from the control block and stop at an empty one.  The counter rests at gridDim.x
    // between launches: the last producer to finish puts it back.
    if (warp == YR_CONSUMERS) {
        // ---- producer: stages tile i+1 (window boxes + the tile's group descriptors) while the consumers work on tile i ----
        if (lane == 0) {
            int tix = blockIdx.x;
            for (int i = 0;; ++i) {
                const int s = i % nbuf;
                if (i >= nbuf) mbar_wait(&ctl.empty[s], ((i / nbuf) - 1) & 1);
                if (tix >= n_all) {                                       // no tile left: an empty one ends the consumers' loop
                    ctl.tile[s].ngroups = 0;
                    mbar_arrive(&ctl.full[s]);
                    break;
                }
                const YRTile t = P.rtiles[tix / nP];
                const int pl = tix % nP;
                ctl.tile[s] = t; ctl.plane[s] = pl; ctl.next_group[s] = YR_CONSUMERS;
                const int nbox = (t.wrows + YR_BOX - 1) / YR_BOX;
                const uint32_t gbytes = (uint32_t)(t.ngroups * sizeof(YRGroup));
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // the buffer was read through the generic proxy
                mbar_expect_tx(&ctl.full[s], (uint32_t)(nbox * YR_BOX * YR_C * sizeof(double)) + gbytes);
                double* win = reinterpret_cast<double*>(smem_raw + CTL + s * wbytes);
                const int prow = pl * P.D.f[t.field].rows_y + t.wlo;
                for (int b = 0; b < nbox; ++b)
                    tma_load_2d(win + (size_t)b * YR_BOX * YR_C, &maps.m[t.field], t.col0, prow + b * YR_BOX, &ctl.full[s]);
                tma_load_1d(&ctl.groups[s][0], P.rgroups + t.g0, gbytes, &ctl.full[s]);
                tix = atomicAdd(P.rcounter, 1);                           // the next tile (this thread has a whole tile's time for the round trip)
            }
            __threadfence();
            if (atomicAdd(P.rcounter + 1, 1) == (int)gridDim.x - 1) {     // every producer has made its last claim
                P.rcounter[0] = (int)gridDim.x;
                P.rcounter[1] = 0;
            }
        }
        return;
    }

    for (int i = 0;; ++i) {
        const int s = i % nbuf;
        mbar_wait_backoff(&ctl.full[s], (i / nbuf) & 1);
        const YRTile& t = ctl.tile[s];
        if (t.ngroups == 0) break;
        const int pl = ctl.plane[s];
        const FieldDev& F = P.D.f[t.field];
        const double* col = reinterpret_cast<const double*>(smem_raw + CTL + s * wbytes) + lane;   // sample of window row r: col[r * YR_C]
        const int x = t.col0 + lane;                             // extended column
        const bool live = x < F.We;
        const int ngroups = t.ngroups, wlo = t.wlo;
        for (int g = warp; g < ngroups;) {
            int nx = 0;
            if (lane == 0) nx = atomicAdd(&ctl.next_group[s], 1);       // the group after this one: asked for now, looked at below
            const YRGroup& G = ctl.groups[s][g];
            const int R = G.nrows, N = G.N;
            const double a = G.a, a4 = G.a4, naN1 = G.naN1, inv_s = G.inv_s;
            const int c = G.j0 + F.Ny_max - wlo;                 // window row of the group's first output row
            const bool lock = G.pad != 0;                        // long group: both sums walk down from the first row (below)
            const int ct = lock ? c : c + R - 1;                 // row the B sum starts at: the group's last row, or its first
            // ---- Horner starts: F at the first row over rows c-N..c, B at the last row over rows ct..ct+N; term i = 4 m + q goes to chain q ----
            double f0 = 0.0, f1 = 0.0, f2 = 0.0, f3 = 0.0, b0 = 0.0, b1 = 0.0, b2 = 0.0, b3 = 0.0;
            {
                int m = N >> 2;
                const int i0 = 4 * m;                            // farthest terms: chains whose index exceeds N have nothing there
                f0 = col[(c - i0) * YR_C]; b0 = col[(ct + i0) * YR_C];
                if (i0 + 1 <= N) { f1 = col[(c - i0 - 1) * YR_C]; b1 = col[(ct + i0 + 1) * YR_C]; }
                if (i0 + 2 <= N) { f2 = col[(c - i0 - 2) * YR_C]; b2 = col[(ct + i0 + 2) * YR_C]; }
                if (i0 + 3 <= N) { f3 = col[(c - i0 - 3) * YR_C]; b3 = col[(ct + i0 + 3) * YR_C]; }
                const double* pf = col + (size_t)(c - i0 + 4) * YR_C;   // term 4(m-1) of chain 0 on the F side
                const double* pb = col + (size_t)(ct + i0 - 4) * YR_C;
                // Software pipeline with two register sets (no copies): the eight samples of step m-1 are requested before the eight FMAs
                // of step m.  Steps come in pairs; an odd step count is settled first.
                auto load8 = [&](double* x) {
                    x[0] = pf[0]; x[1] = pf[-YR_C]; x[2] = pf[-2 * YR_C]; x[3] = pf[-3 * YR_C];
                    x[4] = pb[0]; x[5] = pb[YR_C]; x[6] = pb[2 * YR_C]; x[7] = pb[3 * YR_C];
                    pf += 4 * YR_C; pb -= 4 * YR_C;
                };
                auto fma8 = [&](const double* x) {
                    f0 = __fma_rn(a4, f0, x[0]); b0 = __fma_rn(a4, b0, x[4]);
                    f1 = __fma_rn(a4, f1, x[1]); b1 = __fma_rn(a4, b1, x[5]);
                    f2 = __fma_rn(a4, f2, x[2]); b2 = __fma_rn(a4, b2, x[6]);
                    f3 = __fma_rn(a4, f3, x[3]); b3 = __fma_rn(a4, b3, x[7]);
                };
                double xa[8], xb[8];
                if (m & 1) { load8(xa); fma8(xa); --m; }                 // m steps remain, now even
                if (m > 0) {
                    load8(xa);
                    for (; m > 2; m -= 2) {
                        load8(xb); fma8(xa);
                        load8(xa); fma8(xb);
                    }
                    load8(xb); fma8(xa); fma8(xb);
                }
            }
            double Fc = __fma_rn(a, __fma_rn(a, __fma_rn(a, f3, f2), f1), f0);
            double Bc = __fma_rn(a, __fma_rn(a, __fma_rn(a, b3, b2), b1), b0);
            double* dst = F.r_zs + (size_t)pl * F.ps_zs + (size_t)G.j0 * F.pitch_z + F.zoff + F.yshift + x;   // r_zs interior (df.cpp:377)
            if (lock) {
                // ---- long group (9..YR_JL rows, N >= 25): F and B both start at the first row and walk down in lock step, ----
                //   F_{j+1} = a F_j + (x_{j+1} - a^(N+1) x_{j-N}),      B_{j+1} = B_j / a + (a^N x_{j+1+N} - x_j / a),
                // nothing is kept per row, one dependent FMA per row and sum.  The B walk runs against its stable direction: an error
                // grows by 1/a per row, so the planner bounds the group length by exp(2 pi R / N) <= 8 (R <= 0.33 N): measured
                // deviation from the direct sum <= 6e-15 of the rms at the bound (tests/test_run_form_bound.py; gate 1e-12).
                const double ia = __drcp_rn(a), aN = __dmul_rn(-naN1, ia);
                double xc = col[c * YR_C];
                {
                    const double v = __dmul_rn(inv_s, __dsub_rn(__dadd_rn(Fc, Bc), xc));
                    if (live) dst[0] = v;
                }
                const double* pc = col + (size_t)(c + 1) * YR_C;         // x_{j+1}
                const double* pl_ = pc - (size_t)(N + 1) * YR_C;         // x_{j-N}
                const double* ph = pc + (size_t)N * YR_C;                // x_{j+1+N}
                double* d = dst + F.pitch_z;
#pragma unroll 4
                for (int tt = 1; tt < R; ++tt) {
                    const double xn = *pc, xl = *pl_, xh = *ph;
                    Fc = __fma_rn(a, Fc, __fma_rn(naN1, xl, xn));
                    Bc = __fma_rn(ia, Bc, __fma_rn(aN, xh, -__dmul_rn(ia, xc)));
                    const double v = __dmul_rn(inv_s, __dsub_rn(__dadd_rn(Fc, Bc), xn));
                    if (live) *d = v;
                    xc = xn;
                    pc += YR_C; pl_ += YR_C; ph += YR_C; d += F.pitch_z;
                }
            } else {
            // ---- walks over the group's rows: F up from the first row, B down from the last (both in their stable direction) ----
            double xr[YJ], Fv[YJ], xlo[YJ], xhi[YJ];
#pragma unroll
            for (int tt = 0; tt < YJ; ++tt) {
                xr[tt] = tt < R ? col[(c + tt) * YR_C] : 0.0;
                xlo[tt] = (tt >= 1 && tt < R) ? col[(c + tt - N - 1) * YR_C] : 0.0;      // what the window drops on the way up
                xhi[tt] = tt < R - 1 ? col[(c + tt + N + 1) * YR_C] : 0.0;               // ... on the way down
            }
            Fv[0] = Fc;
#pragma unroll
            for (int tt = 1; tt < YJ; ++tt) {
                if (tt < R) Fc = __fma_rn(naN1, xlo[tt], __fma_rn(a, Fc, xr[tt]));
                Fv[tt] = Fc;
            }
#pragma unroll
            for (int tt = YJ - 1; tt >= 0; --tt) {
                if (tt < R) {
                    if (tt < R - 1) Bc = __fma_rn(naN1, xhi[tt], __fma_rn(a, Bc, xr[tt]));
                    const double v = __dmul_rn(inv_s, __dsub_rn(__dadd_rn(Fv[tt], Bc), xr[tt]));
                    if (live) dst[(size_t)tt * F.pitch_z] = v;
                }
            }
            }
            g = __shfl_sync(0xffffffffu, nx, 0);                        // next group of this tile
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&ctl.empty[s]);
    }
    if (lane == 0) tl_stamp(P.tl, 1);
}

// =================================================================================================
// H2z + H3 + H4 + H5 tuned: z-sweep for row-uniform half-widths with the fused epilogue.
//
// Persistent and warp-independent: every warp is a worker that pulls ITEMS from a global counter -- a (row j, strip of
// 32*ZK columns[, plane of a batch]) as the pair of units u, v run back to back, or as the single unit w; pairs first,
// most expensive first -- through a private double-buffered shared-memory window.  One unit = (row, strip, field):
//   * staging: ONE lane issues two TMA operations per unit -- a 3-D cp.async.bulk.tensor box
//     {ZK doubles, box_lines lines, 1 row} of r_zs with 128/64-byte swizzle (lane l's ZK samples of chunk ch are
//     line l+ch; the swizzle makes the LDS.128 of 8 consecutive lines conflict-free) and a cp.async.bulk of the
//     row's parameter line (recursive form) / padded coefficient vector (direct form) -- both landing on the unit's
//     mbarrier while the previous unit is still computing;
//   * tap loop: lane l owns the ZK consecutive outputs k = c0 + ZK l + (0..ZK-1); recursive evaluation of the
//     exponential window (MODE 1) or register-blocked Toeplitz product (MODE 0), see below;
//   * epilogue per field from registers: filt_old prefetched before the tap loop, blend (H3), Lund scaling (H4),
//     SRA (H5); each of the eight output arrays is written exactly once.  v' = b u_filt + c v_filt (df.cpp:437)
//     needs u's blended value: the u unit leaves its strip in a per-warp shared-memory line buffer (lane-owned
//     layout), the v unit of the same item -- the very next unit of the same warp -- picks it up from there: no
//     flag, no fence, no second trip through global memory;
//   * N2 (opt-in): the running sums of u'^2, v'^2, w'^2, T'^2, rho'^2, u'v' (rms_add, df.cpp:571-582) are
//     accumulated here, where the five values are in registers.
// =================================================================================================
// Cell arithmetic of the fused epilogue with every rounding spelled out, so that the result does not
// depend on which code path (coalesced / direct / partial strip) a cell takes: slabs stay bit-identical
// to the whole plane.
__device__ __forceinline__ double epi_blend(double fo, double sa, double z, double sb) { return __fma_rn(fo, sa, __dmul_rn(z, sb)); }
__device__ __forceinline__ double epi_scale(double rc, double z) { return __dmul_rn(rc, z); }
__device__ __forceinline__ double epi_cross(double rc1, double uf, double o) { return __fma_rn(rc1, uf, o); }


__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}

// Is the strip of a unit handled through the coalesced (piece-major) epilogue?  Its first cell 16-byte aligned and an even number of
// its columns inside the slab (a partial last strip included): every 16-byte piece is then whole and aligned.
__device__ __forceinline__ bool z_strip_coalesced(const ZParams& P, int j, int c0, int plane) {
    const size_t sbase = (size_t)plane * P.D.ps_cells + (size_t)j * P.D.W + c0;
    const int cols = min(32 * P.zk, P.D.W - c0);
    return ((cols & 1) == 0) && ((sbase & 1) == 0) && !(P.debug & 4);
}

__device__ __forceinline__ void z_issue_unit(const ZParams& P, const ZMaps& maps, const int* desc, int plane, void* buf, uint64_t* bar) {
    // called by one lane: up to four TMA operations land everything the unit reads on `bar` --
    //   the window (3-D tensor box of r_zs, swizzled), the row's parameter line / padded coefficient vector,
    //   the row's eight epilogue constants, and (blend steps, coalesced strips) the strip of filt_old --
    // so that the warp itself issues no global load at all: nothing of a unit waits on a memory round trip.
    // The buffer was last touched through the generic proxy (tap-loop reads, transpose scratch):
    // order those accesses before the async-proxy writes of the TMA engine.
    // desc = ZUnit as 8 ints in shared memory: [0] j, [1] c0, [2] f, [3] nchunk, [4] line0, [5] cbytes, [6] coff16
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const int j = desc[0], c0 = desc[1], f = desc[2];
    const uint32_t cbytes = (uint32_t)desc[5];
    const bool fo = !P.S.first_step && z_strip_coalesced(P, j, c0, plane);
    const uint32_t strip_bytes = (uint32_t)min(32 * P.zk, P.D.W - c0) * 8u;      // a partial last strip stages only what exists
    unsigned char* b = reinterpret_cast<unsigned char*>(buf);
    mbar_expect_tx(bar, (uint32_t)P.box_bytes + cbytes + (uint32_t)(ROWC * sizeof(double)) + (fo ? strip_bytes : 0u));
    tma_load_3d(b, &maps.m[f], 0, desc[4], plane * P.D.Ny + j, bar);
    if (P.unit_par) {
        // recursive form: the row's parameter line and its epilogue constants sit side by side in one record per (row, field):
        // one bulk copy instead of two (issuing a TMA operation costs the lone issuing lane ~200 cycles)
        tma_load_1d(b + P.box_bytes, P.unit_par + ((size_t)j * 3 + f) * (16 + ROWC), (16 + ROWC) * sizeof(double), bar);
    } else {
        tma_load_1d(b + P.box_bytes, P.coef_pad + (size_t)desc[6] * 16, cbytes, bar);
        tma_load_1d(b + P.rc_off, P.D.rowc + (size_t)j * ROWC, ROWC * sizeof(double), bar);
    }
    if (fo) tma_load_1d(b + P.fo_off, P.D.f[f].filt_old + (size_t)plane * P.D.ps_cells + (size_t)j * P.D.W + c0, strip_bytes, bar);
}

template <int ZK, int MODE, bool STATS>
__global__ void __maxnreg__(MODE == 1 ? Z_REC_MAXNREG : 168) zsweep_epilogue_kernel(const __grid_constant__ ZMaps maps, const ZParams P) {
    constexpr int LB = ZK * 8;               // bytes per staged line (= one lane's ZK samples): 128 -> SWIZZLE_128B, 64 -> SWIZZLE_64B
    constexpr int PPL = ZK / 2;              // 16-byte pieces per line
    constexpr int UB = 32 * LB;              // bytes of the per-warp u line buffer
    // hardware swizzle of the TMA box: the 16-byte piece index is XORed with address bits [7:9] (128B mode) or [7:8] (64B mode)
    auto swz = [](int line) { return ZK == 16 ? (line & 7) : ((line >> 1) & 3); };
    extern __shared__ __align__(1024) unsigned char zsm_unaligned[];
    // the swizzle is a function of the shared-memory address: put the buffers on a 1 KiB boundary
    unsigned char* zsm_raw = zsm_unaligned + ((1024u - (smem_u32(zsm_unaligned) & 1023u)) & 1023u);
    const PlaneDev& D = P.D;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    // per warp: 2 staging buffers of unit_bytes (window lines + parameter line / coefficient vector) + the u line buffer;
    // then 2 mbarriers and 2 item descriptors (3 units x 8 ints) per warp
    const size_t wbytes = (size_t)2 * P.unit_bytes + UB;
    unsigned char* wbase = zsm_raw + (size_t)warp * wbytes;
    unsigned char* ubuf = wbase + (size_t)2 * P.unit_bytes;
    unsigned char* tail = zsm_raw + (size_t)4 * wbytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(tail) + warp * 2;
    int* descs = reinterpret_cast<int*>(tail + 64) + warp * 32;              // [2 items][<= 2 units][8]
    if (lane == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_fence_init(); }
    // Programmatic dependent launch: this CTA became resident while the y-sweep's last tiles were still running;
    // wait for that grid to complete (and its writes to r_zs to be visible) before the first item is claimed/staged.
    asm volatile("griddepcontrol.wait;" ::: "memory");

    if (lane == 0) tl_stamp(P.tl, 0);
    const long long tstart = (P.debug & 16) ? clock64() : 0;
    unsigned long long gstart = 0;
    if (P.debug & 16) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gstart));
    // Items: first the (u, v) pairs -- two units back to back, v picks u's blended strip up from the warp's line buffer --, then the w
    // units as items of their own (the queue's tail is then one unit long, not three).  Item it of a plane: units [off, off + nun).
    const int nP = D.P, n_total = P.n_items * nP;
    auto item_units = [&](int g) { return (g / nP) < P.n_uv ? 2 : 1; };
    auto item_off = [&](int g) { const int it = g / nP; return it < P.n_uv ? 2 * it : P.n_uv + it; };
    auto fetch_item = [&](int g) {                          // the item's unit descriptors, one int per lane (16 or 8 lanes)
        return lane < 8 * item_units(g) ? __ldg(reinterpret_cast<const int*>(P.units + item_off(g)) + lane) : 0;
    };
    // The first two items of every warp are fixed -- items are sorted most expensive first, warp g of G takes g and G + g -- and
    // the counter starts at 2 G (set by the y-sweep): no warp begins with two round trips to one contended L2 address.
    const int gw = (int)blockIdx.x * 4 + warp, gwarps = (int)gridDim.x * 4;
    int claim = 0;
    int itemA = gw;
    if (itemA >= n_total) return;
    if (lane < 16) descs[lane] = fetch_item(itemA);
    int itemB = gw + gwarps;
    if (itemB < n_total && lane < 16) descs[16 + lane] = fetch_item(itemB);
    __syncwarp();
    int plA = itemA % nP, plB = itemB % nP;                 // plane of a batch (0 for a single plane), kept per item: no division per unit
    int unA = item_units(itemA);
    if (lane == 0) z_issue_unit(P, maps, descs, plA, wbase, &bars[0]);

    // Pipeline per unit n (buffer n & 1, barrier phase (n >> 1) & 1); an item is the three consecutive units u, v, w:
    //   top:    stage unit n+1 (the v unit of this pair, or the first unit of the next item, whose descriptors are already in shared
    //           memory: lane 0, ~650 cycles -- 20 for the proxy fence, ~230 for descriptor reads + expect_tx, ~135 per TMA operation);
    //           at an item's first unit also claim the item after the next (atomic, result not awaited)
    //   middle: tap loop of unit n
    //   bottom: epilogue of unit n; at an item's last unit fetch the descriptors of the newly claimed item into the slot just vacated
    int slot = 0, phase = 0;
    long long pa_wait = 0, pa_taps = 0, pa_epi = 0, pa_unit = 0, pa_top = 0, pa_stage = 0, pa_n = 0;   // DFB_DEBUG_Z & 16: per-warp phase cycles,
                                                                                                       // flushed once at the warp's exit
    for (int n = 0;; ++n) {
        const long long ttop = (P.debug & 16) ? clock64() : 0;
        const int* dcur = descs + slot * 16 + phase * 8;
        const bool last_phase = phase == unA - 1;
        const bool have_next = !last_phase || itemB < n_total;
        if (have_next && lane == 0) {
            unsigned char* nb = wbase + (size_t)((n + 1) & 1) * P.unit_bytes;
            const int* dnext = last_phase ? descs + (slot ^ 1) * 16 : dcur + 8;
            z_issue_unit(P, maps, dnext, last_phase ? plB : plA, nb, &bars[(n + 1) & 1]);
        }
        // (The address is made to look lane-dependent: an atomic add on an address ptxas can prove uniform is turned into a
        //  warp-aggregated atomic whose result is shuffled at once -- the warp would sit out the whole round trip to L2,
        //  ~700 cycles per item, right here.)
        const int zoff = lane & dcur[7];                    // == 0 (ZUnit::pad), but not to the compiler
        if (phase == 0 && itemB < n_total && lane == 0)
            claim = atomicAdd(P.counter + zoff, 1);         // the item after the next: its round trip hides behind the tap loop
        if (P.debug & 16) pa_stage += clock64() - ttop;                                           // staging the next unit (lane 0)

        const int j = dcur[0], c0 = dcur[1], f = dcur[2];
        const int nchunk = __shfl_sync(0xffffffffu, dcur[3], 0);
        const int k0 = c0 + lane * ZK;
        const bool active = k0 < D.W;
        const size_t pbase = (size_t)plA * D.ps_cells;                        // this plane's share of the dense output arrays
        const size_t base = pbase + (size_t)j * D.W + k0;
        const size_t sbase = pbase + (size_t)j * D.W + c0;                    // first cell of the warp's strip
        // whole strip inside the plane and 16-byte aligned: piece-major epilogue (coalesced global access); its filt_old strip was
        // staged with the window
        const bool coalesced = z_strip_coalesced(P, j, c0, plA);
        const FieldDev& F = D.f[f];
        const bool blend = !P.S.first_step;
        const long long tunit = (P.debug & 16) ? clock64() : 0;
        const long long t0 = (P.debug & 16) ? clock64() : 0;
        mbar_wait(&bars[n & 1], (n >> 1) & 1);
        const long long t1 = (P.debug & 16) ? clock64() : 0;
        unsigned char* cbuf = wbase + (size_t)(n & 1) * P.unit_bytes;
        const double* B = reinterpret_cast<const double*>(cbuf + P.box_bytes);
        double acc[ZK];
        if (MODE == 1) {
            // ---- recursive form ----
            // The reference's coefficients are a truncated two-sided exponential, b_i = a^|i| / s, a = exp(-2 pi / N)
            // (df.cpp:168-177), so with F_k = sum_{i=0..N} a^i x_{k-i} and B_k = sum_{i=0..N} a^i x_{k+i}
            //     out_k = (F_k + B_k - x_k) / s,   F_k = a F_{k-1} + x_k - a^(N+1) x_{k-N-1},   B_k = a B_{k+1} + x_k - a^(N+1) x_{k+N+1}.
            // Each lane starts its two recursions with a Horner pass over the N+1 samples before its first / after its last
            // output (the same recurrence with nothing to drop), then walks its ZK outputs: ~2(N + 2 ZK) FMAs per lane instead of
            // ZK (2N+1).  Lane blocks are anchored to plane columns (multiples of 16), so slabs reproduce the whole plane.
            // Every operation is an explicit round-to-nearest intrinsic.  Agreement with the direct sum: ~1e-15 of the rms.
            const double* hdr = reinterpret_cast<const double*>(cbuf + P.box_bytes + dcur[5]) - 16;
            const double a = hdr[0], naN1 = -hdr[1], cn = hdr[2];
            const int Nn = (int)hdr[3], dd = (int)hdr[4];
            const int cl = (dd + Nn) / ZK;                 // the lane's own ZK samples are line cl of its window; cl >= 1
            const int ef = (ZK - 1 + Nn) % ZK;             // last element used of the farthest line (line 2 cl)
            auto load_line = [&](int m, double* x) {
                const int line = lane + m;
                const unsigned char* lp = cbuf + line * LB;
                const int sw = swz(line) << 4;
#pragma unroll
                for (int i = 0; i < ZK / 2; ++i) {
                    const double2 t = *reinterpret_cast<const double2*>(lp + ((i << 4) ^ sw));
                    x[2 * i] = t.x; x[2 * i + 1] = t.y;
                }
            };
            auto lds1 = [&](int p) {                       // sample p of the lane's window (p >= 0)
                const int line = lane + p / ZK, e = p % ZK;
                return *reinterpret_cast<const double*>(cbuf + line * LB + ((((e >> 1) ^ swz(line))) << 4) + (e & 1) * 8);
            };
            // Horner over whole lines as a binary tree: a line's ZK terms are folded pairwise with a, a^2, a^4, (a^8), independent of
            // the running value; only one FMA per line, F <- a^ZK F + T, is on the dependent chain.  A line's sums do not depend
            // on which lane wants them (lanes l and l+1 share all but one line of their windows), so every line of the staged
            // window is read and folded ONCE -- lane l folds line l and, if the window has it, line l + 32 -- and the lanes pick the
            // sums they need with shuffles: 3-4 line reads per lane instead of 2 cl + 1.
            const double a2 = hdr[5], a4 = hdr[6], a8 = hdr[7], aZK = ZK == 16 ? hdr[8] : hdr[7];
            auto fold_c = [&](const double* x, int first) {        // sum_e a^(ZK-1-e) x_e over e >= first (element 0 is the oldest)
                double t[ZK / 2];
#pragma unroll
                for (int q = 0; q < ZK / 2; ++q) t[q] = __fma_rn(a, (2 * q < first) ? 0.0 : x[2 * q], (2 * q + 1 < first) ? 0.0 : x[2 * q + 1]);
#pragma unroll
                for (int q = 0; q < ZK / 4; ++q) t[q] = __fma_rn(a2, t[2 * q], t[2 * q + 1]);
#pragma unroll
                for (int q = 0; q < ZK / 8; ++q) t[q] = __fma_rn(a4, t[2 * q], t[2 * q + 1]);
                if (ZK == 16) t[0] = __fma_rn(a8, t[0], t[1]);
                return t[0];
            };
            auto fold_a = [&](const double* x, int last) {         // sum_e a^e x_e over e <= last (element 0 is the nearest)
                double t[ZK / 2];
#pragma unroll
                for (int q = 0; q < ZK / 2; ++q) t[q] = __fma_rn(a, (2 * q + 1 > last) ? 0.0 : x[2 * q + 1], (2 * q > last) ? 0.0 : x[2 * q]);
#pragma unroll
                for (int q = 0; q < ZK / 4; ++q) t[q] = __fma_rn(a2, t[2 * q + 1], t[2 * q]);
#pragma unroll
                for (int q = 0; q < ZK / 8; ++q) t[q] = __fma_rn(a4, t[2 * q + 1], t[2 * q]);
                if (ZK == 16) t[0] = __fma_rn(a8, t[1], t[0]);
                return t[0];
            };
            double Fc, Bc;
            if (2 * cl <= 32) {
            double TcA, TcM, TaA, TaMA, TcB = 0.0, TaB = 0.0, TaMB = 0.0;
            {
                double x[ZK];
                load_line(0, x);                                   // line `lane`
                TcA = fold_c(x, 0);
                TcM = dd ? fold_c(x, dd) : TcA;                    // as the first line of this lane's own window: elements before d are outside
                TaA = fold_a(x, ZK - 1);
                TaMA = ef < ZK - 1 ? fold_a(x, ef) : TaA;          // as the farthest line of lane (lane - 2 cl)'s window: elements after ef are outside
                if (lane < 2 * cl) {                               // line `lane + 32` exists
                    load_line(32, x);
                    TcB = fold_c(x, 0);
                    TaB = fold_a(x, ZK - 1);
                    TaMB = ef < ZK - 1 ? fold_a(x, ef) : TaB;
                }
            }
            auto pick = [&](double vA, double vB, int line) {      // the sum of window line `line` of this unit (all lanes call together)
                const double fa = __shfl_sync(0xffffffffu, vA, line & 31), fb = __shfl_sync(0xffffffffu, vB, line & 31);
                return line < 32 ? fa : fb;
            };
            Fc = TcM;
            for (int m = 1; m < cl; ++m) Fc = __fma_rn(aZK, Fc, pick(TcA, TcB, lane + m));
            Bc = pick(TaMA, TaMB, lane + 2 * cl);
            for (int m = 1; m < cl; ++m) Bc = __fma_rn(aZK, Bc, pick(TaA, TaB, lane + 2 * cl - m));
            } else {
                // windows longer than 64 lines (N > 256 at ZK = 16): every lane folds its own lines, same sums in the same order
                Fc = 0.0; Bc = 0.0;
                for (int m = 0; m < cl; ++m) {
                    double x[ZK];
                    load_line(m, x);
                    Fc = __fma_rn(aZK, Fc, fold_c(x, m == 0 ? dd : 0));
                    load_line(2 * cl - m, x);
                    Bc = __fma_rn(aZK, Bc, fold_a(x, m == 0 ? ef : ZK - 1));
                }
            }
            double xc[ZK], Fv[ZK];
            load_line(cl, xc);
            // the samples each step drops: x_{k-N-1} = window positions d .. d+ZK-2, x_{k+N+1} = positions d+2N+1 .. d+2N+ZK-1.
            // For even d (even N) they are 16-byte aligned runs: ZK/2 vector loads each; otherwise scalar loads.
            const int pb = dd + Nn + Nn + 1;               // window position of x_{k+N+1} for output 0
            double xo[ZK], xq[ZK];
            if ((dd & 1) == 0) {
                auto load_run = [&](int p0, double* x) {   // ZK samples from even window position p0
#pragma unroll
                    for (int q = 0; q < ZK / 2; ++q) {
                        const int pc_ = (p0 >> 1) + q, line = lane + pc_ / PPL;
                        const double2 t = *reinterpret_cast<const double2*>(cbuf + line * LB + ((((pc_ % PPL) ^ swz(line))) << 4));
                        x[2 * q] = t.x; x[2 * q + 1] = t.y;
                    }
                };
                load_run(dd, xo);                          // xo[kk-1] = x_{k-N-1} of output kk
                load_run(pb - 1, xq);                      // xq[kk+1] = x_{k+N+1} of output kk
            } else {
#pragma unroll
                for (int kk = 0; kk < ZK - 1; ++kk) { xo[kk] = lds1(dd + kk); xq[kk + 1] = lds1(pb + kk); }
            }
#pragma unroll
            for (int kk = 0; kk < ZK; ++kk) {
                Fc = __fma_rn(a, Fc, xc[kk]);
                if (kk > 0) Fc = __fma_rn(naN1, xo[kk - 1], Fc);
                Fv[kk] = Fc;
            }
#pragma unroll
            for (int kk = ZK - 1; kk >= 0; --kk) {
                Bc = __fma_rn(a, Bc, xc[kk]);
                if (kk < ZK - 1) Bc = __fma_rn(naN1, xq[kk + 1], Bc);
                acc[kk] = __dmul_rn(cn, __dsub_rn(__dadd_rn(Fv[kk], Bc), xc[kk]));
            }
        } else {
        double w[2 * ZK - 1];
#pragma unroll
        for (int i = 0; i < ZK; ++i) acc[i] = 0.0;
#pragma unroll
        for (int i = 0; i < ZK - 1; ++i) w[i] = 0.0;
        for (int ch = 0; ch < nchunk; ++ch) {
            const int line = lane + ch;
            const unsigned char* lp = cbuf + line * LB;
            const int sw = swz(line) << 4;               // 16-byte piece i of a line lives at i ^ swz(line)
            double x[ZK];
#pragma unroll
            for (int i = 0; i < ZK / 2; ++i) {
                const double2 t = *reinterpret_cast<const double2*>(lp + ((i << 4) ^ sw));
                x[2 * i] = t.x; x[2 * i + 1] = t.y;
            }
#pragma unroll
            for (int i = 0; i < ZK / 2; ++i) {
                const double2 t = *reinterpret_cast<const double2*>(B + ZK * ch + ZK + 2 * i);
                w[ZK - 1 + 2 * i] = t.x; w[ZK + 2 * i] = t.y;
            }
            // out[kk] += x[q] * b[(ZK ch + q) - kk - d] = x[q] * w[q - kk + ZK - 1]      (df.cpp:397-399)
#pragma unroll
            for (int q = 0; q < ZK; ++q)
#pragma unroll
                for (int kk = 0; kk < ZK; ++kk) acc[kk] = fma(x[q], w[q - kk + ZK - 1], acc[kk]);
#pragma unroll
            for (int i = 0; i < ZK - 1; ++i) w[i] = w[i + ZK];
        }
        }
        __syncwarp();                    // every lane is done with this buffer's window: it becomes the transpose scratch
        // descriptors of the item claimed at this item's first unit: request them now, store them after the epilogue
        int itemC = n_total, dreg = 0;
        if (last_phase && itemB < n_total) {
            itemC = __shfl_sync(0xffffffffu, claim, 0);
            if (itemC < n_total) dreg = fetch_item(itemC);
        }
        const long long t2 = (P.debug & 16) ? clock64() : 0;

        // ---- epilogue of this (strip, field) ----
        {
            const double* rc = reinterpret_cast<const double*>(cbuf + P.rc_off);        // the row's constants, staged with the window
            const double rc_own = rc[f == 0 ? 0 : (f == 1 ? 2 : 3)];
            const double rc0 = rc[0], rc1 = rc[1], rc4 = rc[4], rc5 = rc[5], rc6 = rc[6];
            const double sa = P.S.sa[f], sb = P.S.sb[f];
            double* stats = (STATS && blend) ? P.stats + (size_t)plA * 6 * D.ps_cells + ((size_t)j * D.W) : nullptr;
            if (coalesced) {
                // One transpose per unit: lane l owns line l (its ZK cells) of a 32-line tile; global memory wants 16-byte piece
                // p = lane + 32 m of the strip.  The tap loop's results go through the (swizzled, conflict-free both ways) tile once;
                // everything after is elementwise with per-row constants, so it is done in piece order: filt_old pieces come from the
                // staged strip, every output piece goes straight to its coalesced global address, and u's blended pieces wait for
                // the v unit in the warp's line buffer.
                {
                    const int own = lane * LB, osw = swz(lane) << 4;
#pragma unroll
                    for (int i = 0; i < ZK / 2; ++i)
                        *reinterpret_cast<double2*>(cbuf + own + ((i << 4) ^ osw)) = make_double2(acc[2 * i], acc[2 * i + 1]);
                }
                __syncwarp();
                const double2* fo_t = reinterpret_cast<const double2*>(cbuf + P.fo_off);
                double2* ub = reinterpret_cast<double2*>(ubuf);
                double2* g_fo = reinterpret_cast<double2*>(F.filt_old + sbase);
                double2* g_fl = reinterpret_cast<double2*>(F.fluc + sbase);
                double2* g_T = reinterpret_cast<double2*>(D.T_fluc + sbase);
                double2* g_r = reinterpret_cast<double2*>(D.rho_fluc + sbase);
                auto add2 = [&](int which, int pc, double2 x, double2 y) {              // sums[which] += x * y (mul, then add: df.cpp:575-579)
                    double2* gs = reinterpret_cast<double2*>(stats + (size_t)which * D.ps_cells + c0) + pc;
                    const double2 cur = *gs;
                    *gs = make_double2(__dadd_rn(cur.x, __dmul_rn(x.x, y.x)), __dadd_rn(cur.y, __dmul_rn(x.y, y.y)));
                };
                const int npieces = min(32 * ZK, D.W - c0) >> 1;                            // 16-byte pieces of the strip inside the slab
                auto piece = [&](int m) {
                    const int pc = lane + 32 * m, r = pc / PPL;
                    double2 z = *reinterpret_cast<const double2*>(cbuf + r * LB + ((((pc % PPL) ^ swz(r))) << 4));
                    if (blend) {                                                                 // correlate_fields, df.cpp:415
                        const double2 fo = fo_t[pc];
                        z.x = epi_blend(fo.x, sa, z.x, sb); z.y = epi_blend(fo.y, sa, z.y, sb);
                    }
                    g_fo[pc] = z;                                                                // filt_old <- filt, df.cpp:440-442
                    if (f == 0) ub[pc] = z;                                                      // for the v unit that follows
                    double2 o = make_double2(epi_scale(rc_own, z.x), epi_scale(rc_own, z.y));    // df.cpp:436,438; v.filt term of 437
                    if (f == 1) {
                        const double2 uf = ub[pc];
                        o.x = epi_cross(rc1, uf.x, o.x); o.y = epi_cross(rc1, uf.y, o.y);        // df.cpp:437
                        if (STATS && stats) add2(5, pc, make_double2(epi_scale(rc0, uf.x), epi_scale(rc0, uf.y)), o);   // u' of the same cells
                    }
                    __stcs(g_fl + pc, o);
                    if (STATS && stats) add2(f, pc, o, o);
                    if (f == 0 && blend) {                                                       // get_rho_T_fluc, df.cpp:474-481
                        const double2 t2 = make_double2(__dmul_rn(rc4, o.x), __dmul_rn(rc4, o.y));
                        const double2 Tv = make_double2(__dmul_rn(t2.x, rc5), __dmul_rn(t2.y, rc5));
                        const double2 rv = make_double2(__dmul_rn(-t2.x, rc6), __dmul_rn(-t2.y, rc6));
                        __stcs(g_T + pc, Tv);
                        __stcs(g_r + pc, rv);
                        if (STATS && stats) { add2(3, pc, Tv, Tv); add2(4, pc, rv, rv); }
                    }
                };
                if (npieces == 16 * ZK) {                    // whole strip (the common case): no guard, the eight pieces' loads overlap
#pragma unroll
                    for (int m = 0; m < ZK / 2; ++m) piece(m);
                } else {
#pragma unroll
                    for (int m = 0; m < ZK / 2; ++m)
                        if (lane + 32 * m < npieces) piece(m);
                }
            } else {
                // partial / unaligned strip: every lane walks its own cells in global memory; u's strip waits in the line buffer in
                // lane-owned layout (both units of a pair take the same path)
                double* ul = reinterpret_cast<double*>(ubuf) + lane * ZK;
                if (active) {
                    double* __restrict__ fold = F.filt_old + base;
                    double* __restrict__ fluc = F.fluc + base;
                    double* st = stats ? stats + k0 : nullptr;
#pragma unroll
                    for (int i = 0; i < ZK; ++i) {
                        if (k0 + i < D.W) {
                            double za = acc[i];
                            if (blend) za = epi_blend(fold[i], sa, za, sb);
                            fold[i] = za;
                            if (f == 0) ul[i] = za;
                            double oa = epi_scale(rc_own, za);
                            double ufv = 0.0;
                            if (f == 1) { ufv = ul[i]; oa = epi_cross(rc1, ufv, oa); }
                            fluc[i] = oa;
                            if (STATS && st) {
                                st[(size_t)f * D.ps_cells + i] = __dadd_rn(st[(size_t)f * D.ps_cells + i], __dmul_rn(oa, oa));
                                if (f == 1) st[(size_t)5 * D.ps_cells + i] = __dadd_rn(st[(size_t)5 * D.ps_cells + i], __dmul_rn(epi_scale(rc0, ufv), oa));
                            }
                            if (f == 0 && blend) {
                                const double ta = __dmul_rn(rc4, oa);
                                const double Tv = __dmul_rn(ta, rc5), rv = __dmul_rn(-ta, rc6);
                                D.T_fluc[base + i] = Tv;
                                D.rho_fluc[base + i] = rv;
                                if (STATS && st) {
                                    st[(size_t)3 * D.ps_cells + i] = __dadd_rn(st[(size_t)3 * D.ps_cells + i], __dmul_rn(Tv, Tv));
                                    st[(size_t)4 * D.ps_cells + i] = __dadd_rn(st[(size_t)4 * D.ps_cells + i], __dmul_rn(rv, rv));
                                }
                            }
                        }
                    }
                }
            }
            __syncwarp();                // tile / staged-strip reads are done before the buffer is refilled by the next-but-one unit
        }
        if (P.debug & 16) {
            const long long t3 = clock64();
            pa_wait += t1 - t0; pa_taps += t2 - t1; pa_epi += t3 - t2; pa_unit += t3 - tunit; pa_top += tunit - ttop; pa_n += 1;
        }
        if (!have_next) {
            if (lane == 0) tl_stamp(P.tl, 1);
            if ((P.debug & 16) && lane == 0) {
                atomicAdd(P.prof + 0, (unsigned long long)pa_wait);       // waiting for the staged unit
                atomicAdd(P.prof + 1, (unsigned long long)pa_taps);       // tap loop
                atomicAdd(P.prof + 2, (unsigned long long)pa_epi);        // epilogue
                atomicAdd(P.prof + 3, (unsigned long long)pa_unit);       // whole unit
                atomicAdd(P.prof + 4, (unsigned long long)pa_n);
                atomicAdd(P.prof + 8, (unsigned long long)pa_stage);      // staging the next unit (lane 0)
                atomicAdd(P.prof + 9, (unsigned long long)pa_top);        // top of the loop: staging + claim + addressing
                const unsigned long long life = (unsigned long long)(clock64() - tstart);
                atomicAdd(P.prof + 5, life);
                atomicMax(P.prof + 6, life);
                atomicAdd(P.prof + 7, 1ull);
                // start / end of this warp on the global nanosecond timer: [10] latest start, [11] earliest start (as 2^62 - t),
                // [12] latest end, [13] earliest end (as 2^62 - t), [14] sum of ends - starts
                unsigned long long gend;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gend));
                atomicMax(P.prof + 10, gstart); atomicMax(P.prof + 11, (1ull << 62) - gstart);
                atomicMax(P.prof + 12, gend);   atomicMax(P.prof + 13, (1ull << 62) - gend);
                atomicAdd(P.prof + 14, gend - gstart);
                atomicAdd(P.prof + 16 + min(63, (int)((gend - gstart) / 1000)), 1ull);       // [16..79] histogram of warp lifetimes, 1 us bins
                atomicAdd(P.prof + 80 + min(31, (int)pa_n), 1ull);                           // [80..111] histogram of units per warp
            }
            break;
        }
        if (last_phase) {
            // descriptors of the newly claimed item into the slot this item just vacated
            if (itemC < n_total && lane < 16) descs[slot * 16 + lane] = dreg;
            __syncwarp();
            itemA = itemB; itemB = itemC; slot ^= 1; phase = 0;
            plA = plB; plB = itemC % nP; unA = item_units(itemA);
        } else {
            ++phase;
        }
    }
}

// =================================================================================================
// N2: running sums of squares (rms_add, df.cpp:571-582) + u'v', opt-in.  mul then add, like the reference.
// Separate kernel for the general (one-thread-per-cell) path only; the tuned z-sweep accumulates in its epilogue.
// =================================================================================================
__global__ void __launch_bounds__(256) stats_kernel(const PlaneDev D, double* __restrict__ sums, size_t n) {
    const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;       // cell within the plane
    if (c >= n) return;
    const size_t i = (size_t)blockIdx.y * n + c;                          // cell within the batch's dense arrays
    double* s = sums + (size_t)blockIdx.y * 6 * n + c;                    // [P][6][n]
    const double u = D.f[0].fluc[i], v = D.f[1].fluc[i], w = D.f[2].fluc[i], T = D.T_fluc[i], r = D.rho_fluc[i];
    s[0] = __dadd_rn(s[0], __dmul_rn(u, u));
    s[n] = __dadd_rn(s[n], __dmul_rn(v, v));
    s[2 * n] = __dadd_rn(s[2 * n], __dmul_rn(w, w));
    s[3 * n] = __dadd_rn(s[3 * n], __dmul_rn(T, T));
    s[4 * n] = __dadd_rn(s[4 * n], __dmul_rn(r, r));
    s[5 * n] = __dadd_rn(s[5 * n], __dmul_rn(u, v));
}

// N3: CFD hand-off (ghost cell = mean + fluctuation, the loop a US3D-style plugin runs over its inflow faces, us3d_user.f90:88-113)
__global__ void __launch_bounds__(256) scatter_kernel(const double* __restrict__ field, int n, const int* __restrict__ plane_index,
                                                      const int* __restrict__ dst_index, const double* __restrict__ mean, double scale,
                                                      double* __restrict__ dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = plane_index[i];
    if (c < 0) return;                       // face in another rank's slab (dfb_face_map)
    const int d = dst_index[i];
    const double base = mean ? mean[i] : dst[d];
    dst[d] = __fma_rn(scale, field[c], base);
}

// config 4, destination rank: ONE rank's staged slab [3][Ny][W] -> its columns [k0, k0+W) of the row-major planes [5][Ny][NzG];
// T' and rho' are rebuilt from u' with the row constants, with the epilogue's own roundings (df.cpp:474-481), so the assembled
// plane is bit for bit what the slabs hold.  Launched per source rank as its slab lands (the next rank's transfer runs meanwhile).
__global__ void __launch_bounds__(256) assemble_kernel(const double* __restrict__ slab, double* __restrict__ plane, const double* __restrict__ rowc,
                                                       int Ny, int NzG, int k0, int W, int first_only) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (k >= W) return;
    const size_t cell = (size_t)j * W + k, nslab = (size_t)Ny * W, n = (size_t)Ny * NzG, idx = (size_t)j * NzG + k0 + k;
    const double u = __ldcs(slab + cell), v = __ldcs(slab + nslab + cell), w = __ldcs(slab + 2 * nslab + cell);
    plane[idx] = u; plane[n + idx] = v; plane[2 * n + idx] = w;
    double T = 0.0, rho = 0.0;
    if (!first_only) {
        const double* rc = rowc + (size_t)j * ROWC;
        const double t2 = __dmul_rn(rc[4], u);
        T = __dmul_rn(t2, rc[5]);
        rho = __dmul_rn(-t2, rc[6]);
    }
    plane[3 * n + idx] = T; plane[4 * n + idx] = rho;
}

// config 4, peer-to-peer transport, destination rank: u', v', w' of the columns [k0, k0+W) have just been written into the planes by
// the owning rank's copy engine; rebuild T', rho' of those columns from u' (same roundings as the epilogue, df.cpp:474-481).
__global__ void __launch_bounds__(256) rebuild_kernel(double* __restrict__ plane, const double* __restrict__ rowc, int Ny, int NzG, int k0, int W,
                                                      int first_only) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (k >= W) return;
    const size_t n = (size_t)Ny * NzG, idx = (size_t)j * NzG + k0 + k;
    double T = 0.0, rho = 0.0;
    if (!first_only) {
        const double* rc = rowc + (size_t)j * ROWC;
        const double t2 = __dmul_rn(rc[4], plane[idx]);
        T = __dmul_rn(t2, rc[5]);
        rho = __dmul_rn(-t2, rc[6]);
    }
    plane[3 * n + idx] = T; plane[4 * n + idx] = rho;
}

cudaError_t launch_rebuild(double* plane, const double* rowc, int Ny, int NzG, int k0, int W, int first_only, cudaStream_t st) {
    rebuild_kernel<<<dim3((unsigned)((W + 255) / 256), (unsigned)Ny), 256, 0, st>>>(plane, rowc, Ny, NzG, k0, W, first_only);
    return cudaGetLastError();
}

cudaError_t launch_assemble(const double* slab, double* plane, const double* rowc, int Ny, int NzG, int k0, int W, int first_only, cudaStream_t st) {
    assemble_kernel<<<dim3((unsigned)((W + 255) / 256), (unsigned)Ny), 256, 0, st>>>(slab, plane, rowc, Ny, NzG, k0, W, first_only);
    return cudaGetLastError();
}

cudaError_t launch_scatter(const double* field, int n, const int* plane_index, const int* dst_index, const double* mean, double scale,
                           double* dst, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    scatter_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(field, n, plane_index, dst_index, mean, scale, dst);
    return cudaGetLastError();
}

cudaError_t launch_stats(const PlaneDev& D, double* sums, cudaStream_t st) {
    const size_t n = (size_t)D.Ny * D.W;
    stats_kernel<<<dim3((unsigned)((n + 255) / 256), (unsigned)D.P), 256, 0, st>>>(D, sums, n);
    return cudaGetLastError();
}

// =================================================================================================
// fp64 roofline denominator: dependent-chain-free DFMA issue, 8 chains per thread
// =================================================================================================
__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    const double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;   // defeat dead-code elimination
}

// =================================================================================================
// host launchers
// =================================================================================================
cudaError_t launch_noise(const NoiseParams& P, const PlaneDev& D, cudaStream_t st) {
    if (P.n_arrays == 0) return cudaSuccess;
    cudaGetLastError();        // the status returned below is this launch's, not a leftover of an unrelated earlier call
    int max_seg = 0;
    for (int a = 0; a < P.n_arrays; ++a) max_seg = P.a[a].n_seg > max_seg ? P.a[a].n_seg : max_seg;
    dim3 grid((unsigned)P.chunks, (unsigned)max_seg, (unsigned)(P.n_arrays * D.P));
    noise_kernel<<<grid, NOISE_THREADS, 0, st>>>(P, D);
    return cudaGetLastError();
}

cudaError_t launch_ysweep_simple(const PlaneDev& D, cudaStream_t st) {
    for (int f = 0; f < 3; ++f) {
        dim3 grid((unsigned)((D.f[f].We + 255) / 256), (unsigned)D.Ny, (unsigned)D.P);
        ysweep_simple_kernel<<<grid, 256, 0, st>>>(D, f);
    }
    return cudaGetLastError();
}

cudaError_t launch_zsweep_simple(const PlaneDev& D, const StepConsts& S, cudaStream_t st) {
    dim3 grid((unsigned)((D.W + 255) / 256), (unsigned)D.Ny, (unsigned)D.P);
    zsweep_epilogue_simple_kernel<<<grid, 256, 0, st>>>(D, S);
    return cudaGetLastError();
}

constexpr int Y_RC = 8, Y_NS = 6;

size_t ysweep_smem_bytes() { return sizeof(YSmem<Y_RC, Y_NS>); }
int noise_threads() { return NOISE_THREADS * NOISE_PAIRS; }
int noise_stride_pairs() { return NOISE_THREADS; }
int ysweep_rc() { return Y_RC; }

cudaError_t ysweep_prepare() {
    // One shared-memory carve-out for every kernel of the step: an SM only changes its L1 / shared split when it is idle, so a
    // noise CTA running with the default split beside the sweeps could leave an SM unable to take its third y-sweep CTA (observed
    // as a process-wide bimodal y-sweep time on small planes: 26 or 43 us on the reference default plane).
    {
        cudaError_t e0 = cudaFuncSetAttribute(noise_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e0 != cudaSuccess) return e0;
        e0 = cudaFuncSetAttribute(stats_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e0 != cudaSuccess) return e0;
        e0 = cudaFuncSetAttribute(scatter_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e0 != cudaSuccess) return e0;
    }
    cudaError_t e = cudaFuncSetAttribute(ysweep_tma_kernel<Y_RC, Y_NS, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)sizeof(YSmem<Y_RC, Y_NS, 128>));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(ysweep_tma_kernel<Y_RC, Y_NS, 128>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(ysweep_tma_kernel<Y_RC, Y_NS, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(YSmem<Y_RC, Y_NS, 64>));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(ysweep_tma_kernel<Y_RC, Y_NS, 64>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(ysweep_rec_kernel<Y_RC, Y_NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(YSmem<Y_RC, Y_NS>));
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(ysweep_rec_kernel<Y_RC, Y_NS>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}

static cudaLaunchConfig_t pdl_config(unsigned grid, unsigned block, size_t smem, cudaStream_t st, cudaLaunchAttribute* attr) {
    attr->id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr->val.programmaticStreamSerializationAllowed = 1;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cfg;
}

cudaError_t launch_ysweep_tma(const YMaps& maps, const YParams& P, int n_dense, int n_rec, cudaStream_t st) {
    cudaLaunchAttribute attr;
    YParams Q = P;
    if (n_dense > 0) {
        Q.tile0 = 0;
        Q.n_tiles = n_dense;
        const int n_all = n_dense * P.D.P;
        const int grid = P.resident_grid > 0 ? (n_all < P.resident_grid ? n_all : P.resident_grid) : n_all;
        cudaError_t e;
        if (P.tk == 64) {
            cudaLaunchConfig_t cfg = pdl_config((unsigned)grid, 160, sizeof(YSmem<Y_RC, Y_NS, 64>), st, &attr);
            e = cudaLaunchKernelEx(&cfg, ysweep_tma_kernel<Y_RC, Y_NS, 64>, maps, Q);
        } else {
            cudaLaunchConfig_t cfg = pdl_config((unsigned)grid, 160, sizeof(YSmem<Y_RC, Y_NS, 128>), st, &attr);
            e = cudaLaunchKernelEx(&cfg, ysweep_tma_kernel<Y_RC, Y_NS, 128>, maps, Q);
        }
        if (e != cudaSuccess) return e;
    }
    if (n_rec > 0) {
        Q.tile0 = n_dense;
        cudaLaunchConfig_t cfg = pdl_config((unsigned)(n_rec * P.D.P), 160, sizeof(YSmem<Y_RC, Y_NS>), st, &attr);
        return cudaLaunchKernelEx(&cfg, ysweep_rec_kernel<Y_RC, Y_NS>, maps, Q);
    }
    return cudaSuccess;
}

size_t ysweep_run_smem(int wrows, int nbuf) { return (sizeof(YRSmemCtl) + 127) / 128 * 128 + (size_t)nbuf * wrows * YR_C * sizeof(double); }

cudaError_t ysweep_run_prepare(size_t smem) {
    cudaError_t e = cudaFuncSetAttribute(ysweep_run_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(ysweep_run_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}

cudaError_t launch_ysweep_run(const YRMaps& maps, const YParams& P, cudaStream_t st) {
    cudaLaunchAttribute attr;
    const int n_all = P.n_rtiles * P.D.P;
    cudaLaunchConfig_t cfg = pdl_config((unsigned)(n_all < P.r_grid ? n_all : P.r_grid), 32 * (YR_CONSUMERS + 1), (size_t)P.r_smem, st, &attr);
    return cudaLaunchKernelEx(&cfg, ysweep_run_kernel, maps, P);
}

static const void* zsweep_fn(int zk, int mode, bool stats) {
    if (stats) {
        if (zk == 16) return mode == 1 ? (const void*)zsweep_epilogue_kernel<16, 1, true> : (const void*)zsweep_epilogue_kernel<16, 0, true>;
        return mode == 1 ? (const void*)zsweep_epilogue_kernel<8, 1, true> : (const void*)zsweep_epilogue_kernel<8, 0, true>;
    }
    if (zk == 16) return mode == 1 ? (const void*)zsweep_epilogue_kernel<16, 1, false> : (const void*)zsweep_epilogue_kernel<16, 0, false>;
    return mode == 1 ? (const void*)zsweep_epilogue_kernel<8, 1, false> : (const void*)zsweep_epilogue_kernel<8, 0, false>;
}

cudaError_t zsweep_prepare(int zk, int mode, size_t smem, int* blocks_per_sm) {
    for (int st = 1; st >= 0; --st) {
        const void* fn = zsweep_fn(zk, mode, st != 0);
        cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
    }
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, zsweep_fn(zk, mode, false), 128, smem);
}

cudaError_t launch_zsweep_tuned(const ZMaps& maps, const ZParams& P, cudaStream_t st) {
    cudaLaunchAttribute attr;
    cudaLaunchConfig_t cfg = pdl_config((unsigned)P.nblocks, 128, (size_t)P.smem_bytes, st, &attr);
    static const int pdl = std::getenv("DFB_Z_PDL") ? std::atoi(std::getenv("DFB_Z_PDL")) : 1;
    if (!pdl) cfg.numAttrs = 0;
    void* args[2] = {const_cast<ZMaps*>(&maps), const_cast<ZParams*>(&P)};
    return cudaLaunchKernelExC(&cfg, zsweep_fn(P.zk, P.zmode, P.stats != nullptr), args);
}

cudaError_t launch_dfma_peak(double* out, int blocks, int iters, cudaStream_t st) {
    dfma_peak_kernel<<<blocks, 256, 0, st>>>(out, iters, 1.0000001, 1e-9);
    return cudaGetLastError();
}

}  // namespace dfb
