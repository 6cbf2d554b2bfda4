// Probe: SM partitioning with green contexts (driver API through cudaGetDriverEntryPoint, no -lcuda):
// two partitions, runtime-API kernel launches, event waits across partitions and a captured graph on
// their streams; every kernel records the SM ids it ran on.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o green_probe tools/probes/green_probe.cu && ./green_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <vector>

#define DRV(name) \
    decltype(&name) p_##name = nullptr; \
    { void *f = nullptr; cudaDriverEntryPointQueryResult q; \
      if (cudaGetDriverEntryPoint(#name, &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !f) { printf("no %s\n", #name); return 2; } \
      p_##name = (decltype(&name))f; }
#define CK(x) do { CUresult r_ = (x); if (r_ != CUDA_SUCCESS) { printf("%s -> %d (line %d)\n", #x, (int)r_, __LINE__); return 3; } } while (0)
#define RT(x) do { cudaError_t r_ = (x); if (r_ != cudaSuccess) { printf("%s -> %s (line %d)\n", #x, cudaGetErrorString(r_), __LINE__); return 4; } } while (0)

__global__ void mark(unsigned *seen, int spin) {
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    if (threadIdx.x == 0) atomicOr(&seen[smid >> 5], 1u << (smid & 31));
    long long t0 = clock64();
    while (clock64() - t0 < spin) {}
}

static int count_bits(const unsigned *w, int n, int *lo, int *hi) {
    int c = 0; *lo = 1 << 30; *hi = -1;
    for (int i = 0; i < n * 32; ++i) if ((w[i >> 5] >> (i & 31)) & 1u) { ++c; if (i < *lo) *lo = i; if (i > *hi) *hi = i; }
    return c;
}

int main() {
    RT(cudaSetDevice(0));
    RT(cudaFree(0));
    DRV(cuDeviceGet) DRV(cuDeviceGetDevResource) DRV(cuDevSmResourceSplitByCount) DRV(cuDevResourceGenerateDesc)
    DRV(cuGreenCtxCreate) DRV(cuGreenCtxStreamCreate) DRV(cuGreenCtxDestroy)
    CUdevice dev; CK(p_cuDeviceGet(&dev, 0));
    CUdevResource all; CK(p_cuDeviceGetDevResource(dev, &all, CU_DEV_RESOURCE_TYPE_SM));
    printf("device SMs: %u\n", all.sm.smCount);
    CUdevResource part[2], rest; unsigned n = 1;
    CK(p_cuDevSmResourceSplitByCount(&part[0], &n, &all, &rest, 0, 64));
    printf("split: groups %u, first %u SMs, remaining %u SMs\n", n, part[0].sm.smCount, rest.sm.smCount);
    CUdevResourceDesc d0, d1; CK(p_cuDevResourceGenerateDesc(&d0, &part[0], 1)); CK(p_cuDevResourceGenerateDesc(&d1, &rest, 1));
    CUgreenCtx g0, g1; CK(p_cuGreenCtxCreate(&g0, d0, dev, CU_GREEN_CTX_DEFAULT_STREAM)); CK(p_cuGreenCtxCreate(&g1, d1, dev, CU_GREEN_CTX_DEFAULT_STREAM));
    CUstream s0, s1; CK(p_cuGreenCtxStreamCreate(&s0, g0, CU_STREAM_NON_BLOCKING, 0)); CK(p_cuGreenCtxStreamCreate(&s1, g1, CU_STREAM_NON_BLOCKING, 0));
    unsigned *seen; RT(cudaMalloc(&seen, 3 * 8 * 4)); RT(cudaMemset(seen, 0, 3 * 8 * 4));
    // plain runtime launches on the two partitions
    mark<<<592, 64, 0, (cudaStream_t)s0>>>(seen, 20000);
    RT(cudaGetLastError());
    mark<<<592, 64, 0, (cudaStream_t)s1>>>(seen + 8, 20000);
    RT(cudaGetLastError());
    // event from partition 0, waited on by partition 1
    cudaEvent_t ev; RT(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    RT(cudaEventRecord(ev, (cudaStream_t)s0));
    RT(cudaStreamWaitEvent((cudaStream_t)s1, ev, 0));
    // a captured graph on partition 0's stream, launched there
    cudaGraph_t graph; cudaGraphExec_t exec;
    RT(cudaStreamBeginCapture((cudaStream_t)s0, cudaStreamCaptureModeRelaxed));
    mark<<<592, 64, 0, (cudaStream_t)s0>>>(seen + 16, 20000);
    RT(cudaStreamEndCapture((cudaStream_t)s0, &graph));
    RT(cudaGraphInstantiate(&exec, graph, 0));
    RT(cudaGraphLaunch(exec, (cudaStream_t)s0));
    RT(cudaDeviceSynchronize());
    unsigned h[24]; RT(cudaMemcpy(h, seen, sizeof(h), cudaMemcpyDeviceToHost));
    const char *names[3] = {"partition 0 launch", "partition 1 launch", "graph on partition 0"};
    for (int k = 0; k < 3; ++k) { int lo, hi; int c = count_bits(h + 8 * k, 8, &lo, &hi); printf("%s: %d SMs used, ids %d..%d\n", names[k], c, lo, hi); }
    unsigned overlap = 0; for (int i = 0; i < 8; ++i) overlap |= h[i] & h[8 + i];
    printf("partitions overlap: %s; graph stays in partition 0: %s\n", overlap ? "YES" : "no",
           [&] { for (int i = 0; i < 8; ++i) if (h[16 + i] & ~h[i]) return "NO"; return "yes"; }());
    // concurrency: a long kernel on partition 0 must not delay a short one on partition 1
    cudaEvent_t a, b; RT(cudaEventCreate(&a)); RT(cudaEventCreate(&b));
    mark<<<64, 512, 0, (cudaStream_t)s0>>>(seen, 4000000);
    RT(cudaEventRecord(a, (cudaStream_t)s1));
    mark<<<84 * 4, 256, 0, (cudaStream_t)s1>>>(seen + 8, 20000);
    RT(cudaEventRecord(b, (cudaStream_t)s1));
    RT(cudaDeviceSynchronize());
    float ms; RT(cudaEventElapsedTime(&ms, a, b));
    printf("short kernel on partition 1 beside a 2 ms kernel on partition 0: %.3f ms\n", ms);
    printf("PROBE OK\n");
    return 0;
}
