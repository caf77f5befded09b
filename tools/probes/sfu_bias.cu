// Probe: systematic error of sum cos / sum sin evaluated with sin.approx / cos.approx (FMUL.RZ + MUFU) over a uniform grid of
// angles t_i = (i + 0.5) / n turns, whose exact sums are 0.  Variants: the 2 pi constant (previous float / nearest / next
// float / nearest + compensation term, x = fma(t, c, t * lo)) and centring the argument to [-1/2, 1/2] turns first.
// Build: nvcc -arch=sm_100a -O3 -o tools/probes/sfu_bias tools/probes/sfu_bias.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <math.h>
__global__ void k(int n, float c, float lo, int centre, double* out)
{
    double sc = 0, ss = 0, sa = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float t = ((float)i + 0.5f) / (float)n;
        if (centre) t = t - rintf(t);
        float s, co;
        __sincosf(fmaf(t, c, t * lo), &s, &co);
        double es, ec;
        sincospi(2.0 * (double)t, &es, &ec);
        sc += (double)co - ec; ss += (double)s - es; sa += fabs((double)co - ec);
    }
    atomicAdd(out, sc); atomicAdd(out + 1, ss); atomicAdd(out + 2, sa);
}
int main()
{
    const int n = 1 << 24;   // exactly representable grid in fp32
    double* d; cudaMalloc(&d, 24);
    const float c0 = 6.283185307179586f, c1 = nextafterf(c0, 10.f), cm = nextafterf(c0, 0.f);
    const float cs[6] = {cm, c0, c1, c0, c0, c0}, los[6] = {0.f, 0.f, 0.f, 2.5e-7f, 3.1e-7f, 3.7e-7f};
    for (int centre = 0; centre < 2; ++centre)
        for (int v = 0; v < 6; ++v) {
            cudaMemset(d, 0, 24);
            k<<<1024, 256>>>(n, cs[v], los[v], centre, d);
            double h[3]; cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
            printf("centre %d  2pi const %.9g + %.2g : mean error cos %+.3e  sin %+.3e  mean |err cos| %.3e   (2^-24 = %.3e)\n", centre, cs[v], los[v], h[0] / n, h[1] / n, h[2] / n, ldexp(1.0, -24));
        }
    return 0;
}
