// minimal warp-specialised pattern: warps 0-1 meet at a named barrier with an explicit thread count, warp 2 does not
// take part; everybody meets at __syncthreads().  Well-defined PTX (bar.sync a, b); used to see what synccheck says.
#include <cstdio>
__global__ void k(int* out) {
    const int warp = threadIdx.x >> 5;
    __shared__ int s[64];
    if (warp < 2) {
        s[threadIdx.x] = threadIdx.x;
        asm volatile("bar.sync 1, 64;" ::: "memory");
        out[threadIdx.x] = s[63 - threadIdx.x];
        asm volatile("bar.sync 1, 64;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 95) out[64] = 1;
}
int main() {
    int* d; cudaMalloc(&d, 65 * sizeof(int));
    k<<<1, 96>>>(d);
    int h[65]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%d %d %d err=%s\n", h[0], h[63], h[64], cudaGetErrorString(cudaGetLastError()));
    return 0;
}
