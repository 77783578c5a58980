#include <cuda_runtime.h>
#include <cstdio>
__global__ void body(int* counter, cudaGraphConditionalHandle h) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    int c = atomicAdd(counter, 1);
    if (c >= 9) cudaGraphSetConditional(h, 0);
  }
}
int main() {
  int* d; cudaMalloc(&d, 4); cudaMemset(d, 0, 4);
  cudaGraph_t g; cudaGraphCreate(&g, 0);
  cudaGraphConditionalHandle h;
  cudaGraphConditionalHandleCreate(&h, g, 1, cudaGraphCondAssignDefault);
  cudaGraphNodeParams p = {};
  p.type = cudaGraphNodeTypeConditional;
  p.conditional.handle = h; p.conditional.type = cudaGraphCondTypeWhile; p.conditional.size = 1;
  cudaGraphNode_t node;
  cudaError_t e = cudaGraphAddNode(&node, g, nullptr, 0, &p);
  printf("add cond: %s\n", cudaGetErrorString(e));
  cudaGraph_t bodyg = p.conditional.phGraph_out[0];
  cudaKernelNodeParams kp = {};
  void* args[] = {&d, &h};
  kp.func = (void*)body; kp.gridDim = dim3(2); kp.blockDim = dim3(32); kp.kernelParams = args;
  cudaGraphNode_t kn;
  e = cudaGraphAddKernelNode(&kn, bodyg, nullptr, 0, &kp);
  printf("add kernel: %s\n", cudaGetErrorString(e));
  cudaGraphExec_t ex;
  e = cudaGraphInstantiate(&ex, g, 0);
  printf("inst: %s\n", cudaGetErrorString(e));
  cudaStream_t s; cudaStreamCreate(&s);
  for (int rep = 0; rep < 2; ++rep) {
    cudaMemsetAsync(d, 0, 4, s);
    e = cudaGraphLaunch(ex, s);
    cudaStreamSynchronize(s);
    int hv; cudaMemcpy(&hv, d, 4, cudaMemcpyDeviceToHost);
    printf("launch: %s counter=%d\n", cudaGetErrorString(e), hv);
  }
}
