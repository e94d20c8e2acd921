// render.cuh — renderSDF and simulate8bit on the device (/root/reference/core/render-sdf.cpp, bitmap-interpolation.hpp:11-27).
//
// One thread per OUTPUT texel: bilinear sample of the distance field at the texel centre (clamped like the reference's
// interpolate()), median for multi-channel -> single-channel, then either a hard threshold or the inverse distance mapping.
// Purely HBM-bound (each output texel reads 4 neighbouring field texels that neighbouring threads share through L1/L2).
// Arithmetic follows the reference's types: the two horizontal mixes are rounded to float, weights stay in fp64.
#pragma once

#include "kernels.cuh"

namespace b200 {

struct RenderParams {
    const float *sdf;
    float *out;
    int sw, sh, ow, oh;
    double scaleX, scaleY;          // sdf.width/output.width, sdf.height/output.height
    int hard;                       // sdfPxRange.lower == sdfPxRange.upper: threshold instead of mapping
    double mapScale, mapTranslate;  // DistanceMapping::inverse(range scaled to the output size)
    float threshold, bias;          // sdThreshold, .5f - sdThreshold
};

__device__ __forceinline__ int clamp_index(int n, int b) { return n >= 0 && n <= b ? n : (n > 0)*b; }               // clamp(n, b)
__device__ __forceinline__ double clamp_to(double n, double b) { return n >= 0 && n <= b ? n : double(n > 0)*b; }
__device__ __forceinline__ float mix_fd(float a, float b, double w) { return float((1.0-w)*a+w*b); }                  // mix<float, double>

template <int NI, int NO>
__global__ void __launch_bounds__(CTA_THREADS) k6_render_sdf(RenderParams R) {
    const size_t total = size_t(R.ow)*R.oh;
    for (size_t i = size_t(blockIdx.x)*blockDim.x+threadIdx.x; i < total; i += size_t(gridDim.x)*blockDim.x) {
        const int x = int(i%size_t(R.ow)), y = int(i/size_t(R.ow));
        double px = clamp_to(R.scaleX*(x+.5), double(R.sw)), py = clamp_to(R.scaleY*(y+.5), double(R.sh));
        px -= .5;
        py -= .5;
        int l = int(floor(px)), b = int(floor(py));
        int r = l+1, t = b+1;
        const double lr = px-l, bt = py-b;
        l = clamp_index(l, R.sw-1); r = clamp_index(r, R.sw-1);
        b = clamp_index(b, R.sh-1); t = clamp_index(t, R.sh-1);
        const float *lb = R.sdf+(size_t(b)*R.sw+l)*NI, *rb = R.sdf+(size_t(b)*R.sw+r)*NI;
        const float *lt = R.sdf+(size_t(t)*R.sw+l)*NI, *rt = R.sdf+(size_t(t)*R.sw+r)*NI;
        float sd[NI];
        #pragma unroll
        for (int k = 0; k < NI; ++k)
            sd[k] = mix_fd(mix_fd(lb[k], rb[k], lr), mix_fd(lt[k], rt[k], lr), bt);
        float v[NI];
        int nv = NI;
        if (NO == 1 && NI >= 3) {
            v[0] = median3f(sd[0], sd[1], sd[2]);
            nv = 1;
        } else {
            #pragma unroll
            for (int k = 0; k < NI; ++k) v[k] = sd[k];
        }
        #pragma unroll
        for (int k = 0; k < NI; ++k)
            if (k < nv) {
                if (R.hard)
                    v[k] = float(v[k] >= R.threshold);
                else {
                    const double m = R.mapScale*(double(v[k]+R.bias)+R.mapTranslate)+.5;        // distVal: clamp(mapping(dist)+.5)
                    v[k] = float(m >= 0 && m <= 1 ? m : double(m > 0));
                }
            }
        float *o = R.out+i*NO;
        #pragma unroll
        for (int k = 0; k < NO; ++k)
            o[k] = (NO == 3 && NI == 1) ? v[0] : v[k < NI ? k : 0];
    }
}

// simulate8bit: pixelByteToFloat(pixelFloatToByte(x)) (pixel-conversion.hpp:8-14), in place on n floats
__global__ void __launch_bounds__(CTA_THREADS) k6_simulate_8bit(float *pixels, size_t n) {
    for (size_t i = size_t(blockIdx.x)*blockDim.x+threadIdx.x; i < n; i += size_t(gridDim.x)*blockDim.x) {
        const float x = pixels[i];
        const float c = x >= 0.f && x <= 1.f ? x : float(x > 0.f);
        const unsigned char b = (unsigned char) ~int(255.5f-255.f*c);
        pixels[i] = (1.f/255.f)*float(b);
    }
}

}  // namespace b200
