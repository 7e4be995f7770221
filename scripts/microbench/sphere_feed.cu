// Microbenchmark (design experiment, not product): how fast can a warp be fed the sphere list?
//   A: broadcast LDS.128 from shared memory (the shipped kernels) — 512 B per warp per sphere through the
//      128 B/clk/SM crossbar = 4 clk per sphere per SM
//   B: __constant__ memory through the uniform datapath (LDCU -> uniform registers -> FFMA UR operands)
// Both run the 8-instruction conservative filter of rt_trace.cuh over n spheres, `iters` rays per lane.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o sphere_feed sphere_feed.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
struct F4 { float x, y, z, w; };
__constant__ F4 c_hot[4000];

template <bool CONST>
__global__ void __launch_bounds__(256, 4) feed(float* out, const F4* gsph, int n, int iters)
{
    extern __shared__ F4 s_hot[];
    if (!CONST) { for (int i = threadIdx.x; i < n; i += blockDim.x) s_hot[i] = gsph[i]; __syncthreads(); }
    int tid = blockIdx.x * blockDim.x + threadIdx.x;
    float ox = 0.01f * (tid & 63), oy = 0.02f * ((tid >> 6) & 63), oz = 0.5f;
    float dx = 0.6f, dy = 0.0f, dz = -0.8f;
    float acc = 0.f;
    for (int it = 0; it < iters; ++it) {
        float od = ox * dx + oy * dy + oz * dz, nk = -(ox * ox + oy * oy + oz * oz);
        float px = 2.f * ox, py = 2.f * oy, pz = 2.f * oz;
        for (int i = 0; i < n; i += 8) {
            float v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                F4 s = CONST ? c_hot[i + k] : s_hot[i + k];
                float hb = fmaf(-s.x, dx, fmaf(-s.y, dy, fmaf(-s.z, dz, od)));
                float a  = fmaf(s.x, px, fmaf(s.y, py, fmaf(s.z, pz, nk)));
                v[k] = fmaf(hb, hb, a) - s.w;
            }
            float m = v[0];
#pragma unroll
            for (int k = 1; k < 8; ++k) m = fmaxf(m, v[k]);
            if (m >= 0.f) {
#pragma unroll
                for (int k = 0; k < 8; ++k) if (v[k] >= 0.f) acc += sqrtf(v[k]) + (float)(i + k);
            }
        }
        ox += 1e-6f + acc * 1e-12f; dx -= acc * 1e-12f;
    }
    out[tid] = acc;
}

int main(int argc, char** argv)
{
    int n = argc > 1 ? atoi(argv[1]) : 1000, iters = argc > 2 ? atoi(argv[2]) : 200;
    n = (n + 7) / 8 * 8;
    std::vector<F4> h(n);
    for (int i = 0; i < n; ++i) { h[i] = {10.f + (i % 37) * 0.5f, -0.3f, -5.f - (i / 37) * 0.5f, 1e3f}; }   // all miss (w large)
    F4* d; cudaMalloc(&d, n * sizeof(F4)); cudaMemcpy(d, h.data(), n * sizeof(F4), cudaMemcpyHostToDevice);
    cudaMemcpyToSymbol(c_hot, h.data(), n * sizeof(F4));
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int grid = sms * 4, block = 256;
    float* out; cudaMalloc(&out, grid * block * sizeof(float));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int mode = 0; mode < 2; ++mode)
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            if (mode) feed<true><<<grid, block>>>(out, d, n, iters);
            else      feed<false><<<grid, block, n * sizeof(F4)>>>(out, d, n, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double tests = (double)grid * block * iters * n;
            printf("%s n=%d rep%d: %.3f ms  %.1f G lane-tests/s  %.2f clk/warp-sphere/SM (at 1.965 GHz)  err=%s\n",
                   mode ? "const/uniform" : "smem LDS.128", n, rep, ms, tests / ms / 1e6,
                   1.965e9 * (ms * 1e-3) / (tests / 32 / sms), cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}
