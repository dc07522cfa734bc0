// TEST INFRASTRUCTURE — C-ABI launchers around the reference's OWN cubemap-filter kernels (the nvdiffrec
// `renderutils` plugin vendored by the reference at pbr/renderutils/c_src/cubemap.cu). The reference source is
// compiled from where it lies under /root/reference (it is #included below through -I, nothing is copied);
// the host side here only restates what torch_bindings.cpp:740-889 does around each launch: fill the `Tensor`
// descriptors (dims / contiguous strides / _dims), zero the gradient buffer, launch over an [N, N, 6] grid.
// Only tests/, bench.py's reference legs and __graft_entry__.smoke() may load the resulting library.
#include <cstdint>
#include <cfloat>
#include <cstdio>
#include <cuda_runtime.h>
#include "cubemap.cu"   // reference: pbr/renderutils/c_src/cubemap.cu (kernels are TU-local symbols)

namespace {

Tensor nhwc(const void* p, int n, int h, int w, int c, dim3 grid)
{
    Tensor t;
    t.val = const_cast<void*>(p);
    t.d_val = nullptr;
    t.dims[0] = n; t.dims[1] = h; t.dims[2] = w; t.dims[3] = c;
    t.strides[0] = h * w * c; t.strides[1] = w * c; t.strides[2] = c; t.strides[3] = 1;
    t._dims[0] = grid.z; t._dims[1] = grid.y; t._dims[2] = grid.x; t._dims[3] = c;
    t.fp16 = false;
    return t;
}

// torch_bindings.cpp uses getLaunchBlockSize(8, 8, dims); the results do not depend on the block shape
void shape(int N, dim3& grid, dim3& block, dim3& launch)
{
    grid = dim3(N, N, 6);
    block = dim3(8, 8, 1);
    launch = dim3((N + 7) / 8, (N + 7) / 8, 6);
}

int done(const char* what)
{
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { fprintf(stderr, "ref_cubemap %s: %s\n", what, cudaGetErrorString(e)); return (int)e; }
    return 0;
}

}  // namespace

extern "C" {

// cubemap, out: float [6,N,N,3]
int ref_diffuse_cubemap_fwd(int N, const float* cubemap, float* out)
{
    dim3 g, b, l; shape(N, g, b, l);
    DiffuseCubemapKernelParams p;
    p.gridSize = g;
    p.cubemap = nhwc(cubemap, 6, N, N, 3, g);
    p.out = nhwc(out, 6, N, N, 3, g);
    DiffuseCubemapFwdKernel<<<l, b>>>(p);
    return done("diffuse fwd");
}

// grad: float [6,N,N,3] upstream; cubemap_grad: float [6,N,N,3], zeroed here like torch::zeros in the binding
int ref_diffuse_cubemap_bwd(int N, const float* cubemap, const float* grad, float* cubemap_grad)
{
    dim3 g, b, l; shape(N, g, b, l);
    cudaMemset(cubemap_grad, 0, sizeof(float) * 6 * N * N * 3);
    DiffuseCubemapKernelParams p;
    p.gridSize = g;
    p.cubemap = nhwc(cubemap, 6, N, N, 3, g);
    p.out = nhwc(grad, 6, N, N, 3, g);
    p.cubemap.d_val = cubemap_grad;
    DiffuseCubemapBwdKernel<<<l, b>>>(p);
    return done("diffuse bwd");
}

// bounds: float [6,N,N,24]
int ref_specular_bounds(int N, float costheta_cutoff, float* bounds)
{
    dim3 g, b, l; shape(N, g, b, l);
    SpecularBoundsKernelParams p;
    p.costheta_cutoff = costheta_cutoff;
    p.gridSize = g;
    p.out = nhwc(bounds, 6, N, N, 24, g);
    SpecularBoundsKernel<<<l, b>>>(p);
    return done("specular bounds");
}

// out: float [6,N,N,4] = (sum col*w, wsum); the division happens in Python in the reference (ops.py:456)
int ref_specular_cubemap_fwd(int N, const float* cubemap, const float* bounds, float roughness, float costheta_cutoff,
                             float* out)
{
    dim3 g, b, l; shape(N, g, b, l);
    SpecularCubemapKernelParams p;
    p.roughness = roughness;
    p.costheta_cutoff = costheta_cutoff;
    p.gridSize = g;
    p.cubemap = nhwc(cubemap, 6, N, N, 3, g);
    p.bounds = nhwc(bounds, 6, N, N, 24, g);
    p.out = nhwc(out, 6, N, N, 4, g);
    SpecularCubemapFwdKernel<<<l, b>>>(p);
    return done("specular fwd");
}

// grad: float [6,N,N,4] (the kernel reads channels 0..2 only); cubemap_grad [6,N,N,3] zeroed here
int ref_specular_cubemap_bwd(int N, const float* cubemap, const float* bounds, const float* grad, float roughness,
                             float costheta_cutoff, float* cubemap_grad)
{
    dim3 g, b, l; shape(N, g, b, l);
    cudaMemset(cubemap_grad, 0, sizeof(float) * 6 * N * N * 3);
    SpecularCubemapKernelParams p;
    p.roughness = roughness;
    p.costheta_cutoff = costheta_cutoff;
    p.gridSize = g;
    p.cubemap = nhwc(cubemap, 6, N, N, 3, g);
    p.bounds = nhwc(bounds, 6, N, N, 24, g);
    p.out = nhwc(grad, 6, N, N, 4, g);
    p.cubemap.d_val = cubemap_grad;
    SpecularCubemapBwdKernel<<<l, b>>>(p);
    return done("specular bwd");
}

}  // extern "C"
