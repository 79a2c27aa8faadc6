// tools/dfma_probe.cu -- development microbenchmarks for the fp64 pipe (not part of the product).
//   A: x = fma(x, a, b)                    (2 loop-invariant operands: best case for the register file)
//   B: acc[i][j] += c[i] * x[j], 8x4       (the y-sweep's register tile, operands all in registers)
//   C: B + 6 LDS.128 per 32 DFMA           (the y-sweep's exact instruction mix, smem operands)
//   D: Toeplitz 8x8 window                 (the z-sweep's mix)
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(128) kA(double* out, int iters, double a, double b) {
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = fma(x[i], a, b);
    double s = 0; for (int i = 0; i < 8; ++i) s += x[i];
    if (s == 1.2345) out[0] = s;
}

__global__ void __launch_bounds__(128) kB(double* out, const double* in, int iters) {
    double acc[8][4], c[8], x[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i] = in[i]; for (int j = 0; j < 4; ++j) acc[i][j] = 0; }
#pragma unroll
    for (int j = 0; j < 4; ++j) x[j] = in[8 + j + threadIdx.x];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fma(c[i], x[j], acc[i][j]);
        // keep c/x "live and changing" cheaply so nothing is hoisted: rotate through a register swap
        double t = c[0];
#pragma unroll
        for (int i = 0; i < 7; ++i) c[i] = c[i + 1];
        c[7] = t;
    }
    double s = 0;
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += acc[i][j];
    if (s == 1.2345) out[0] = s;
}

__global__ void __launch_bounds__(128) kC(double* out, const double* in, int iters) {
    __shared__ __align__(16) double sx[8][4][128];     // [row][warp][col]
    __shared__ __align__(16) double sc[8][8];
    for (int i = threadIdx.x; i < 8 * 4 * 128; i += 128) (&sx[0][0][0])[i] = in[i % 64];
    if (threadIdx.x < 64) (&sc[0][0])[threadIdx.x] = in[threadIdx.x];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) acc[i][j] = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const double2 xa = *reinterpret_cast<const double2*>(&sx[r][warp][2 * lane]);
            const double2 xb = *reinterpret_cast<const double2*>(&sx[r][warp][64 + 2 * lane]);
#pragma unroll
            for (int jj = 0; jj < 8; jj += 2) {
                const double2 cc = *reinterpret_cast<const double2*>(&sc[r][jj]);
                acc[jj][0] = fma(cc.x, xa.x, acc[jj][0]); acc[jj][1] = fma(cc.x, xa.y, acc[jj][1]);
                acc[jj][2] = fma(cc.x, xb.x, acc[jj][2]); acc[jj][3] = fma(cc.x, xb.y, acc[jj][3]);
                acc[jj + 1][0] = fma(cc.y, xa.x, acc[jj + 1][0]); acc[jj + 1][1] = fma(cc.y, xa.y, acc[jj + 1][1]);
                acc[jj + 1][2] = fma(cc.y, xb.x, acc[jj + 1][2]); acc[jj + 1][3] = fma(cc.y, xb.y, acc[jj + 1][3]);
            }
        }
        __syncwarp();
    }
    double s = 0;
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += acc[i][j];
    if (s == 1.2345) out[0] = s;
}

__global__ void __launch_bounds__(128) kD(double* out, const double* in, int iters) {
    __shared__ __align__(16) double sx[160 * 10 + 64];
    __shared__ __align__(16) double sb[512];
    for (int i = threadIdx.x; i < 160 * 10 + 64; i += 128) sx[i] = in[i % 64];
    for (int i = threadIdx.x; i < 512; i += 128) sb[i] = in[i % 64];
    __syncthreads();
    const double* xs = sx + threadIdx.x * 10;
    double acc[8], w[15];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0;
#pragma unroll
    for (int i = 0; i < 7; ++i) w[i] = 0;
    for (int it = 0; it < iters; ++it) {
        for (int ch = 0; ch < 32; ++ch) {
            double x[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) { const double2 t = *reinterpret_cast<const double2*>(xs + ch * 10 + 2 * i); x[2 * i] = t.x; x[2 * i + 1] = t.y; }
#pragma unroll
            for (int i = 0; i < 4; ++i) { const double2 t = *reinterpret_cast<const double2*>(sb + 8 * ch + 8 + 2 * i); w[7 + 2 * i] = t.x; w[8 + 2 * i] = t.y; }
#pragma unroll
            for (int q = 0; q < 8; ++q)
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) acc[kk] = fma(x[q], w[q - kk + 7], acc[kk]);
#pragma unroll
            for (int i = 0; i < 7; ++i) w[i] = w[i + 8];
        }
    }
    double s = 0;
    for (int i = 0; i < 8; ++i) s += acc[i];
    if (s == 1.2345) out[0] = s;
}


// E: y-mix with 16 rows x 4 cols per thread (64 accumulators): 2 sample + 8 coefficient LDS.128 per 64 DFMA
__global__ void __launch_bounds__(128) kE(double* out, const double* in, int iters) {
    __shared__ __align__(16) double sx[8][4][128];
    __shared__ __align__(16) double sc[8][16];
    for (int i = threadIdx.x; i < 8 * 4 * 128; i += 128) (&sx[0][0][0])[i] = in[i % 64];
    (&sc[0][0])[threadIdx.x] = in[threadIdx.x % 64];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double acc[16][4];
#pragma unroll
    for (int i = 0; i < 16; ++i) for (int j = 0; j < 4; ++j) acc[i][j] = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const double2 xa = *reinterpret_cast<const double2*>(&sx[r][warp][2 * lane]);
            const double2 xb = *reinterpret_cast<const double2*>(&sx[r][warp][64 + 2 * lane]);
#pragma unroll
            for (int jj = 0; jj < 16; jj += 2) {
                const double2 cc = *reinterpret_cast<const double2*>(&sc[r][jj]);
                acc[jj][0] = fma(cc.x, xa.x, acc[jj][0]); acc[jj][1] = fma(cc.x, xa.y, acc[jj][1]);
                acc[jj][2] = fma(cc.x, xb.x, acc[jj][2]); acc[jj][3] = fma(cc.x, xb.y, acc[jj][3]);
                acc[jj + 1][0] = fma(cc.y, xa.x, acc[jj + 1][0]); acc[jj + 1][1] = fma(cc.y, xa.y, acc[jj + 1][1]);
                acc[jj + 1][2] = fma(cc.y, xb.x, acc[jj + 1][2]); acc[jj + 1][3] = fma(cc.y, xb.y, acc[jj + 1][3]);
            }
        }
        __syncwarp();
    }
    double s = 0;
    for (int i = 0; i < 16; ++i) for (int j = 0; j < 4; ++j) s += acc[i][j];
    if (s == 1.2345) out[0] = s;
}

// F: Toeplitz 16x16 window (z-mix with 16 outputs per thread): 8 + 8 LDS.128 per 256 DFMA
__global__ void __launch_bounds__(128) kF(double* out, const double* in, int iters) {
    __shared__ __align__(16) double sx[160 * 18 + 64];
    __shared__ __align__(16) double sb[640];
    for (int i = threadIdx.x; i < 160 * 18 + 64; i += 128) sx[i] = in[i % 64];
    for (int i = threadIdx.x; i < 640; i += 128) sb[i] = in[i % 64];
    __syncthreads();
    const double* xs = sx + threadIdx.x * 18;
    double acc[16], w[31];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = 0;
#pragma unroll
    for (int i = 0; i < 15; ++i) w[i] = 0;
    for (int it = 0; it < iters; ++it) {
        for (int ch = 0; ch < 16; ++ch) {
            double x[16];
#pragma unroll
            for (int i = 0; i < 8; ++i) { const double2 t = *reinterpret_cast<const double2*>(xs + ch * 18 + 2 * i); x[2 * i] = t.x; x[2 * i + 1] = t.y; }
#pragma unroll
            for (int i = 0; i < 8; ++i) { const double2 t = *reinterpret_cast<const double2*>(sb + 16 * ch + 16 + 2 * i); w[15 + 2 * i] = t.x; w[16 + 2 * i] = t.y; }
#pragma unroll
            for (int q = 0; q < 16; ++q)
#pragma unroll
                for (int kk = 0; kk < 16; ++kk) acc[kk] = fma(x[q], w[q - kk + 15], acc[kk]);
#pragma unroll
            for (int i = 0; i < 15; ++i) w[i] = w[i + 16];
        }
    }
    double s = 0;
    for (int i = 0; i < 16; ++i) s += acc[i];
    if (s == 1.2345) out[0] = s;
}

// G: Toeplitz y-mix: 8 rows x 4 cols per thread, uniform N over the 8 rows: 1 coefficient per input row (sliding window of 8)
__global__ void __launch_bounds__(128) kG(double* out, const double* in, int iters) {
    __shared__ __align__(16) double sx[8][4][128];
    __shared__ __align__(16) double sb[64];
    for (int i = threadIdx.x; i < 8 * 4 * 128; i += 128) (&sx[0][0][0])[i] = in[i % 64];
    if (threadIdx.x < 64) sb[threadIdx.x] = in[threadIdx.x];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double acc[8][4], w[15];
#pragma unroll
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) acc[i][j] = 0;
#pragma unroll
    for (int i = 0; i < 7; ++i) w[i] = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i) { const double2 t = *reinterpret_cast<const double2*>(sb + (it & 3) * 8 + 2 * i); w[7 + 2 * i] = t.x; w[8 + 2 * i] = t.y; }
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const double2 xa = *reinterpret_cast<const double2*>(&sx[r][warp][2 * lane]);
            const double2 xb = *reinterpret_cast<const double2*>(&sx[r][warp][64 + 2 * lane]);
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                const double c = w[r - jj + 7];
                acc[jj][0] = fma(c, xa.x, acc[jj][0]); acc[jj][1] = fma(c, xa.y, acc[jj][1]);
                acc[jj][2] = fma(c, xb.x, acc[jj][2]); acc[jj][3] = fma(c, xb.y, acc[jj][3]);
            }
        }
#pragma unroll
        for (int i = 0; i < 7; ++i) w[i] = w[i + 8];
    }
    double s = 0;
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += acc[i][j];
    if (s == 1.2345) out[0] = s;
}

// H: kF with the coefficient address made (trivially) lane-dependent: ptxas can no longer prove it warp-uniform,
//    so the coefficients stay in vector registers -- the operand mix of the production z-sweep
__global__ void __launch_bounds__(128) kH(double* out, const double* in, int iters, int zero) {
    __shared__ __align__(16) double sx[160 * 18 + 64];
    __shared__ __align__(16) double sb[640];
    for (int i = threadIdx.x; i < 160 * 18 + 64; i += 128) sx[i] = in[i % 64];
    for (int i = threadIdx.x; i < 640; i += 128) sb[i] = in[i % 64];
    __syncthreads();
    const double* xs = sx + threadIdx.x * 18;
    const double* bb = sb + threadIdx.x * zero;      // always + 0, but not provably uniform
    double acc[16], w[31];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = 0;
#pragma unroll
    for (int i = 0; i < 15; ++i) w[i] = 0;
    for (int it = 0; it < iters; ++it) {
        for (int ch = 0; ch < 16; ++ch) {
            double x[16];
#pragma unroll
            for (int i = 0; i < 8; ++i) { const double2 t = *reinterpret_cast<const double2*>(xs + ch * 18 + 2 * i); x[2 * i] = t.x; x[2 * i + 1] = t.y; }
#pragma unroll
            for (int i = 0; i < 8; ++i) { const double2 t = *reinterpret_cast<const double2*>(bb + 16 * ch + 16 + 2 * i); w[15 + 2 * i] = t.x; w[16 + 2 * i] = t.y; }
#pragma unroll
            for (int q = 0; q < 16; ++q)
#pragma unroll
                for (int kk = 0; kk < 16; ++kk) acc[kk] = fma(x[q], w[q - kk + 15], acc[kk]);
#pragma unroll
            for (int i = 0; i < 15; ++i) w[i] = w[i + 16];
        }
    }
    double s = 0;
    for (int i = 0; i < 16; ++i) s += acc[i];
    if (s == 1.2345) out[0] = s;
}

template <class F>
static double timeit(F f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) { cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); if (r) best = ms < best ? ms : best; }
    return best * 1e-3;
}

int main() {
    double *out, *in; cudaMalloc(&out, 1024); cudaMalloc(&in, 8192 * 8);
    cudaMemset(in, 0, 8192 * 8);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sm = p.multiProcessorCount;
    for (int bps : {1, 2, 3, 4}) {
        const int blocks = sm * bps, iters = 2000;
        double tA = timeit([&] { kA<<<blocks, 128>>>(out, iters * 8, 1.0000001, 1e-9); });
        double tB = timeit([&] { kB<<<blocks, 128>>>(out, in, iters * 8); });
        double tC = timeit([&] { kC<<<blocks, 128>>>(out, in, iters); });
        double tD = timeit([&] { kD<<<blocks, 128>>>(out, in, iters / 4); });
        double tE = timeit([&] { kE<<<blocks, 128>>>(out, in, iters / 2); });
        double tF = timeit([&] { kF<<<blocks, 128>>>(out, in, iters / 8); });
        double tG = timeit([&] { kG<<<blocks, 128>>>(out, in, iters); });
        double tH = timeit([&] { kH<<<blocks, 128>>>(out, in, iters / 8, 0); });
        double fE = 2.0 * blocks * 128.0 * (iters / 2) * 512, fF = 2.0 * blocks * 128.0 * (iters / 8) * 16 * 256, fG = 2.0 * blocks * 128.0 * iters * 256;
        double fA = 2.0 * blocks * 128.0 * iters * 8 * 32, fB = 2.0 * blocks * 128.0 * iters * 8 * 32, fC = 2.0 * blocks * 128.0 * iters * 256, fD = 2.0 * blocks * 128.0 * (iters / 4) * 32 * 64;
        printf("blocks/SM=%d (warps/SMSP=%d)  A %.2f  B %.2f  C %.2f  D %.2f  E %.2f  F %.2f  G %.2f  H %.2f TFLOP/s\n", bps, bps, fA / tA / 1e12, fB / tB / 1e12, fC / tC / 1e12, fD / tD / 1e12, fE / tE / 1e12, fF / tF / 1e12, fG / tG / 1e12, fF / tH / 1e12);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(e));
    return 0;
}
