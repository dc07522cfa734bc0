// Probe: this toolkit's device expf / division / sqrtf against torch's CUDA kernels. Reads N floats from in.bin,
// writes expf(x), 1/(1+expf(-x)), x / 3.7f.  nvcc -gencode arch=compute_100a,code=sm_100a -o expf_probe expf_probe.cu
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
__global__ void k(int n, const float* x, float* e, float* s, float* q)
{
    int i = blockIdx.x * 256 + threadIdx.x;
    if (i < n) { e[i] = expf(x[i]); s[i] = 1.0f / (1.0f + expf(-x[i])); q[i] = x[i] / 3.7f; }
}
int main(int argc, char** argv)
{
    FILE* f = fopen(argv[1], "rb");
    fseek(f, 0, SEEK_END); long bytes = ftell(f); fseek(f, 0, SEEK_SET);
    int n = bytes / 4;
    std::vector<float> h(n), o(3 * (size_t)n);
    fread(h.data(), 4, n, f); fclose(f);
    float *d, *r;
    cudaMalloc(&d, bytes); cudaMalloc(&r, 3 * bytes);
    cudaMemcpy(d, h.data(), bytes, cudaMemcpyHostToDevice);
    k<<<(n + 255) / 256, 256>>>(n, d, r, r + n, r + 2 * (size_t)n);
    cudaMemcpy(o.data(), r, 3 * bytes, cudaMemcpyDeviceToHost);
    f = fopen(argv[2], "wb"); fwrite(o.data(), 4, 3 * (size_t)n, f); fclose(f);
    return cudaDeviceSynchronize() != cudaSuccess;
}
