#include <cuda_runtime.h>
#include <cstdio>
__global__ void k_body(int *c, cudaGraphConditionalHandle h) { int v = ++*c; cudaGraphSetConditional(h, v < 5); }
int main() {
    int *d; cudaMalloc(&d, 4); cudaMemset(d, 0, 4);
    cudaGraph_t g; cudaGraphCreate(&g, 0);
    cudaGraphConditionalHandle h; cudaGraphConditionalHandleCreate(&h, g, 1, cudaGraphCondAssignDefault);
    cudaGraphNodeParams cp = {cudaGraphNodeTypeConditional};
    cp.conditional.handle = h; cp.conditional.type = cudaGraphCondTypeWhile; cp.conditional.size = 1;
    cudaGraphNode_t node; cudaError_t e = cudaGraphAddNode(&node, g, nullptr, 0, &cp);
    printf("add %d\n", (int)e);
    cudaGraph_t body = cp.conditional.phGraph_out[0];
    cudaKernelNodeParams kp = {}; void *args[] = {&d, &h};
    kp.func = (void *)k_body; kp.gridDim = dim3(1); kp.blockDim = dim3(1); kp.kernelParams = args;
    cudaGraphNode_t kn; e = cudaGraphAddKernelNode(&kn, body, nullptr, 0, &kp); printf("k %d\n", (int)e);
    cudaGraphExec_t ex; e = cudaGraphInstantiate(&ex, g, 0); printf("inst %d\n", (int)e);
    cudaGraphLaunch(ex, 0); cudaDeviceSynchronize();
    int hst; cudaMemcpy(&hst, d, 4, cudaMemcpyDeviceToHost); printf("count %d\n", hst);
}
