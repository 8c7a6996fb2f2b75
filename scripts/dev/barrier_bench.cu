// Microbenchmark: cost of a grid-wide barrier inside a persistent kernel vs a chain of dependent (PDL) kernel launches
// in a CUDA graph. nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/bb scripts/dev/barrier_bench.cu
#include <cuda_runtime.h>
#include <cstdio>
__device__ __forceinline__ void grid_barrier(unsigned* ctr, unsigned target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
    unsigned v;
    long long t0 = clock64();
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
    } while (v < target && clock64() - t0 < 200000000LL);
  }
  __syncthreads();
}
__global__ void bar_kernel(unsigned* ctr, int n, float* sink) {
  float acc = 0;
  for (int i = 0; i < n; ++i) {
    grid_barrier(ctr, (i + 1) * gridDim.x);
    acc += i;
  }
  if (acc < 0) sink[0] = acc;
}
__global__ void empty_kernel(float* sink) {
  asm volatile("griddepcontrol.launch_dependents;");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (sink == nullptr) return;
  if (threadIdx.x == 0 && blockIdx.x == 0) sink[1] += 1.f;
}
int main() {
  unsigned* ctr; float* sink;
  cudaMalloc(&ctr, 4); cudaMalloc(&sink, 64); cudaMemset(sink, 0, 64);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  for (int grid : {sms, 128, 64}) for (int threads : {128, 320}) {
    const int n = 2000;
    for (int rep = 0; rep < 2; ++rep) {
      cudaMemset(ctr, 0, 4);
      cudaEventRecord(e0);
      bar_kernel<<<grid, threads>>>(ctr, n, sink);
      cudaEventRecord(e1); cudaDeviceSynchronize();
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("grid barrier: grid=%d threads=%d  %.3f us/barrier  (%s)\n", grid, threads, ms * 1e3 / n, cudaGetErrorString(cudaGetLastError()));
  }
  // chain of PDL kernels in a graph
  cudaStream_t s; cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
  for (int pdl = 0; pdl < 2; ++pdl) for (int grid : {128, 256}) {
    cudaGraph_t g; cudaGraphExec_t ge;
    const int n = 1000;
    cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
    for (int i = 0; i < n; ++i) {
      cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.stream = s;
      cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = at; cfg.numAttrs = pdl;
      cudaLaunchKernelEx(&cfg, empty_kernel, sink);
    }
    cudaStreamEndCapture(s, &g); cudaGraphInstantiate(&ge, g, 0);
    cudaGraphLaunch(ge, s); cudaStreamSynchronize(s);
    cudaEventRecord(e0, s); cudaGraphLaunch(ge, s); cudaEventRecord(e1, s); cudaStreamSynchronize(s);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("kernel chain in graph: pdl=%d grid=%d  %.3f us/kernel (%s)\n", pdl, grid, ms * 1e3 / n, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
