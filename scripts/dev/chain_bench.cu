// Microbenchmark: cost per kernel of a DEPENDENT chain inside a CUDA graph (each kernel reads what its predecessor
// wrote), with programmatic dependent launch, for a few kernel shapes. nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
template <int WORK>
__global__ void dep_kernel(const float* __restrict__ in, float* __restrict__ out, const float4* __restrict__ w, int nw4) {
  extern __shared__ float sm[];
  float4 acc = make_float4(0, 0, 0, 0);
  // "weights": independent of the predecessor, may be fetched before the wait
  if (WORK) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nw4; i += gridDim.x * blockDim.x) {
      float4 v = __ldg(w + i);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  }
  asm volatile("griddepcontrol.launch_dependents;");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  float x = __ldcg(in + (threadIdx.x & 31));
  sm[threadIdx.x] = x + acc.x + acc.y + acc.z + acc.w;
  __syncthreads();
  if (blockIdx.x == 0 && threadIdx.x < 32) out[threadIdx.x] = sm[threadIdx.x] * 0.5f + 1.f;
}
int main() {
  float *a, *b; float4* w;
  const int nw4 = 6 * 1024 * 1024 / 16;  // 6 MB of "weights" per kernel
  cudaMalloc(&a, 256); cudaMalloc(&b, 256); cudaMemset(a, 0, 256); cudaMemset(b, 0, 256);
  const int n = 480;
  cudaMalloc(&w, (size_t)n * nw4 * 16); cudaMemset(w, 0, (size_t)n * nw4 * 16);
  cudaStream_t s; cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  struct Cfg { int grid, threads, smem, work, pdl; };
  std::vector<Cfg> cfgs = {{1, 32, 0, 0, 1}, {148, 256, 0, 0, 1}, {148, 256, 0, 0, 0}, {296, 256, 100 * 1024, 0, 1}, {148, 256, 0, 1, 1}, {296, 256, 0, 1, 1},
                           {592, 256, 0, 1, 1}, {592, 256, 0, 1, 0}};
  cudaFuncSetAttribute(dep_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(dep_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (auto c : cfgs) {
    cudaGraph_t g; cudaGraphExec_t ge;
    cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
    for (int i = 0; i < n; ++i) {
      cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(c.grid); cfg.blockDim = dim3(c.threads); cfg.stream = s;
      cfg.dynamicSmemBytes = c.smem < 1024 ? c.threads * 4 : c.smem;
      cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = at; cfg.numAttrs = c.pdl;
      const float* in = (i & 1) ? b : a; float* out = (i & 1) ? a : b;
      if (c.work) cudaLaunchKernelEx(&cfg, dep_kernel<1>, in, out, (const float4*)(w + (size_t)i * nw4), nw4);
      else cudaLaunchKernelEx(&cfg, dep_kernel<0>, in, out, (const float4*)w, nw4);
    }
    cudaStreamEndCapture(s, &g); cudaGraphInstantiate(&ge, g, 0);
    cudaGraphLaunch(ge, s); cudaStreamSynchronize(s);
    cudaEventRecord(e0, s); cudaGraphLaunch(ge, s); cudaEventRecord(e1, s); cudaStreamSynchronize(s);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("dependent chain: grid=%4d threads=%3d smem=%6d stream_6MB=%d pdl=%d : %.2f us/kernel (%s)\n", c.grid, c.threads, c.smem, c.work, c.pdl,
           ms * 1e3 / n, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
