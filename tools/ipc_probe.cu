// tools/ipc_probe.cu -- development probe: bandwidth of writes from GPU 1 into GPU 0's memory opened through CUDA IPC in another
// process: (a) cudaMemcpyAsync (copy engine), (b) cudaMemcpy2DAsync (strided rows), (c) a plain copy kernel with a small footprint.
//   nvcc -arch=sm_100a -o ipc_probe ipc_probe.cu && ./ipc_probe
#include <cstdio>
#include <cstdlib>
#include <unistd.h>
#include <sys/wait.h>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
__global__ void copyk(double2* __restrict__ dst, const double2* __restrict__ src, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}
int main() {
    int p1[2], p2[2];
    if (pipe(p1) || pipe(p2)) return 1;
    const size_t n = 64ull << 20;          // doubles: 512 MB
    pid_t pid = fork();
    if (pid == 0) {                        // child: GPU 1, writes into GPU 0's buffer
        cudaIpcMemHandle_t hd;
        if (read(p1[0], &hd, sizeof(hd)) != (ssize_t)sizeof(hd)) return 2;
        CK(cudaSetDevice(1));
        cudaError_t pe = cudaDeviceEnablePeerAccess(0, 0);
        printf("enable peer access: %s\n", cudaGetErrorString(pe));
        void* peer = nullptr;
        CK(cudaIpcOpenMemHandle(&peer, hd, cudaIpcMemLazyEnablePeerAccess));
        double* src; CK(cudaMalloc(&src, n * 8)); CK(cudaMemset(src, 1, n * 8));
        cudaStream_t st; CK(cudaStreamCreate(&st));
        cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
        float ms;
        for (int rep = 0; rep < 2; ++rep) {
            CK(cudaEventRecord(a, st));
            for (int i = 0; i < 5; ++i) CK(cudaMemcpyAsync(peer, src, n * 8, cudaMemcpyDeviceToDevice, st));
            CK(cudaEventRecord(b, st)); CK(cudaEventSynchronize(b)); CK(cudaEventElapsedTime(&ms, a, b));
        }
        printf("cudaMemcpyAsync peer:   %.1f GB/s\n", 5 * n * 8 / (ms * 1e-3) / 1e9);
        for (int rep = 0; rep < 2; ++rep) {
            CK(cudaEventRecord(a, st));
            for (int i = 0; i < 5; ++i) CK(cudaMemcpy2DAsync(peer, 8192 * 8, src, 4096 * 8, 4096 * 8, n / 8192, cudaMemcpyDeviceToDevice, st));
            CK(cudaEventRecord(b, st)); CK(cudaEventSynchronize(b)); CK(cudaEventElapsedTime(&ms, a, b));
        }
        printf("cudaMemcpy2DAsync peer: %.1f GB/s (rows of 32 KB into a 64 KB pitch)\n", 5 * (n / 2) * 8 / (ms * 1e-3) / 1e9);
        for (int grid : {16, 64, 148, 592}) {
            for (int rep = 0; rep < 2; ++rep) {
                CK(cudaEventRecord(a, st));
                for (int i = 0; i < 5; ++i) copyk<<<grid, 256, 0, st>>>((double2*)peer, (const double2*)src, n / 2);
                CK(cudaEventRecord(b, st)); CK(cudaEventSynchronize(b)); CK(cudaEventElapsedTime(&ms, a, b));
            }
            printf("copy kernel %3d x 256:   %.1f GB/s\n", grid, 5 * n * 8 / (ms * 1e-3) / 1e9);
        }
        CK(cudaIpcCloseMemHandle(peer));
        char ok = 1; if (write(p2[1], &ok, 1) != 1) return 3;
        return 0;
    }
    CK(cudaSetDevice(0));
    double* buf; CK(cudaMalloc(&buf, n * 8));
    cudaIpcMemHandle_t hd; CK(cudaIpcGetMemHandle(&hd, buf));
    if (write(p1[1], &hd, sizeof(hd)) != (ssize_t)sizeof(hd)) return 4;
    char ok = 0; if (read(p2[0], &ok, 1) != 1) printf("child failed\n");
    int st = 0; waitpid(pid, &st, 0);
    return 0;
}
