// Measures the cost of a kernel boundary inside a CUDA graph on this GPU, with and without programmatic dependent launch.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/graph_gap_probe scripts/probes/graph_gap_probe.cu && gpurun_out/graph_gap_probe
// A chain of N dependent kernels (each: grid CTAs x 256 threads, `work` dependent FMAs per thread, reads what its
// predecessor wrote) is captured into a graph and replayed; time/N - pure work time = boundary cost.
#include <cuda_runtime.h>
#include <stdio.h>
#include <vector>

template <bool PDL>
__global__ void link_kernel(const float* __restrict__ in, float* __restrict__ out, int work) {
  if (PDL) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // let the next grid's CTAs get scheduled early
  if (PDL) asm volatile("griddepcontrol.wait;" ::: "memory");                // ...but read only after the predecessor has completed
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float v = in[i];
  for (int k = 0; k < work; ++k) v = fmaf(v, 1.0000001f, 1e-7f);
  out[i] = v;
}

static float run(bool pdl, int n_kernels, int grid, int work, int reps) {
  float *a, *b;
  cudaMalloc(&a, grid * 256 * 4);
  cudaMalloc(&b, grid * 256 * 4);
  cudaMemset(a, 0, grid * 256 * 4);
  cudaStream_t st;
  cudaStreamCreate(&st);
  cudaGraph_t g;
  cudaGraphExec_t ge;
  cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
  for (int k = 0; k < n_kernels; ++k) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(256);
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl ? 1 : 0;
    const float* in = (k & 1) ? b : a;
    float* out = (k & 1) ? a : b;
    if (pdl) cudaLaunchKernelEx(&cfg, link_kernel<true>, in, out, work);
    else cudaLaunchKernelEx(&cfg, link_kernel<false>, in, out, work);
  }
  cudaError_t e = cudaStreamEndCapture(st, &g);
  if (e != cudaSuccess) { printf("capture failed: %s\n", cudaGetErrorString(e)); return -1.f; }
  e = cudaGraphInstantiate(&ge, g, 0);
  if (e != cudaSuccess) { printf("instantiate failed: %s\n", cudaGetErrorString(e)); return -1.f; }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int r = 0; r < 3; ++r) cudaGraphLaunch(ge, st);
  cudaStreamSynchronize(st);
  cudaEventRecord(e0, st);
  for (int r = 0; r < reps; ++r) cudaGraphLaunch(ge, st);
  cudaEventRecord(e1, st);
  cudaStreamSynchronize(st);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  e = cudaGetLastError();
  if (e != cudaSuccess) printf("error: %s\n", cudaGetErrorString(e));
  cudaGraphExecDestroy(ge);
  cudaGraphDestroy(g);
  cudaFree(a);
  cudaFree(b);
  return 1e3f * ms / reps / n_kernels;   // us per kernel
}

int main() {
  const int N = 2000;
  printf("# us per kernel inside a %d-kernel dependent chain captured in one CUDA graph\n", N);
  printf("# grid  work   plain_us   pdl_us\n");
  const int grids[] = {1, 148, 296, 1184};
  const int works[] = {0, 2000, 20000};
  for (int g : grids)
    for (int w : works) {
      const float p = run(false, N, g, w, 5), q = run(true, N, g, w, 5);
      printf("%6d %6d %9.3f %9.3f\n", g, w, p, q);
    }
  return 0;
}
