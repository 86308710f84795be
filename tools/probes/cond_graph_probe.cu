#include <cuda_runtime.h>
#include <cstdio>
__global__ void body(int *cnt, cudaGraphConditionalHandle h) {
  int c = atomicAdd(cnt, 1);
  cudaGraphSetConditional(h, c + 1 < 5 ? 1u : 0u);
}
__global__ void setc(cudaGraphConditionalHandle h, unsigned v) { cudaGraphSetConditional(h, v); }
int main() {
  int *cnt; cudaMalloc(&cnt, 4); cudaMemset(cnt, 0, 4);
  cudaStream_t st; cudaStreamCreate(&st);
  cudaGraph_t g; cudaGraphCreate(&g, 0);
  cudaGraphConditionalHandle h;
  printf("%d\n", cudaGraphConditionalHandleCreate(&h, g, 0, 0));
  cudaStreamBeginCaptureToGraph(st, g, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed);
  setc<<<1,1,0,st>>>(h, 1);
  cudaStreamCaptureStatus status; const cudaGraphNode_t *deps; size_t ndeps;
  cudaStreamGetCaptureInfo_v2(st, &status, nullptr, nullptr, &deps, &ndeps);
  cudaGraphNodeParams p = {}; p.type = cudaGraphNodeTypeConditional;
  p.conditional.handle = h; p.conditional.type = cudaGraphCondTypeWhile; p.conditional.size = 1;
  cudaGraphNode_t node;
  printf("add %d\n", cudaGraphAddNode(&node, g, deps, ndeps, &p));
  cudaGraph_t bg = p.conditional.phGraph_out[0];
  cudaStreamUpdateCaptureDependencies(st, &node, 1, cudaStreamSetCaptureDependencies);
  // capture body into bg with another stream
  cudaStream_t st2; cudaStreamCreate(&st2);
  cudaStreamBeginCaptureToGraph(st2, bg, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed);
  body<<<1,1,0,st2>>>(cnt, h);
  cudaGraph_t tmp; printf("endbody %d\n", cudaStreamEndCapture(st2, &tmp));
  setc<<<1,1,0,st>>>(h, 0);
  printf("end %d\n", cudaStreamEndCapture(st, &tmp));
  cudaGraphExec_t ex; printf("inst %d\n", cudaGraphInstantiate(&ex, g, 0));
  cudaGraphLaunch(ex, st); cudaStreamSynchronize(st);
  int hc; cudaMemcpy(&hc, cnt, 4, cudaMemcpyDeviceToHost); printf("cnt=%d err=%d\n", hc, cudaGetLastError());
  cudaGraphLaunch(ex, st); cudaStreamSynchronize(st);
  cudaMemcpy(&hc, cnt, 4, cudaMemcpyDeviceToHost); printf("cnt=%d err=%d\n", hc, cudaGetLastError());
}
