// synth.cu -- synthetic multi-station captures generated on the device (bench / tests), in place of
// the SDRplay stream callback (receiver/capt_sched.c:105-148): 100 baud +-85 Hz continuous-phase
// FSK ('B' = +85 Hz, 'Y' = -85 Hz) at a per-stream channel offset, plus AWGN from a counter-based
// generator keyed by (seed, stream, absolute sample index) so any block of any stream can be
// regenerated independently; values are rounded to the int16 grid the radio delivers.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <vector>

#include "../../include/navtex_b200.h"

namespace {

constexpr int kSamplesPerBit = NVX_FS_HZ / 100;

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

struct SynthArgs {
    const int8_t* sign;        // per bit: +1 ('B', +85 Hz) / -1 ('Y', -85 Hz)
    const int* cum;            // per bit: sum of the signs of all earlier bits of the same stream
    const long long* bit_off;  // [S+1]
    const float* offset_hz;
    const long long* start;    // emission start, samples
    const float* amplitude;
    const float* sigma;
    unsigned long long seed;
    long long t0, n;
    int streams;
    float2* out;
};

__global__ void synth_kernel(const SynthArgs a) {
    const int s = blockIdx.y;
    const long long nb = a.bit_off[s + 1] - a.bit_off[s];
    const double off = (double)a.offset_hz[s];
    const long long start = a.start[s];
    const float amp = a.amplitude[s], sigma = a.sigma[s];
    float2* row = a.out + (size_t)s * a.n;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < a.n; k += (long long)gridDim.x * blockDim.x) {
        const long long t = a.t0 + k;
        float vi = 0.f, vq = 0.f;
        const long long tau = t - start;
        if (tau >= 0) {
            const long long b = tau / kSamplesPerBit;
            if (b < nb) {
                const long long r = tau - b * kSamplesPerBit;
                const size_t bi = (size_t)(a.bit_off[s] + b);
                const double tone = 85.0 * ((double)kSamplesPerBit * (double)a.cum[bi] + (double)a.sign[bi] * (double)(r + 1));
                double cyc = (off * (double)(tau + 1) + tone) / (double)NVX_FS_HZ;
                cyc -= floor(cyc);
                double sn, cs;
                sincospi(2.0 * cyc, &sn, &cs);
                vi = amp * (float)cs;
                vq = amp * (float)sn;
            }
        }
        if (sigma > 0.f) {
            const uint64_t h = mix64(a.seed ^ mix64(((uint64_t)(unsigned)s << 40) ^ (uint64_t)t));
            const float u1 = ((float)(uint32_t)(h >> 32) + 1.0f) * 2.3283064365386963e-10f;
            const float u2 = (float)(uint32_t)h * 2.3283064365386963e-10f;
            const float rad = sigma * sqrtf(-2.0f * __logf(u1));
            float sn, cs;
            __sincosf(6.283185307179586f * u2, &sn, &cs);
            vi += rad * cs;
            vq += rad * sn;
        }
        vi = fminf(fmaxf(rintf(vi), -32768.f), 32767.f);
        vq = fminf(fmaxf(rintf(vq), -32768.f), 32767.f);
        row[k] = make_float2(vi, vq);
    }
}

}  // namespace

extern "C" int nvx_synth_fill_device(int device, const nvx_synth_desc* d, int n_streams, long long t0, long long n, void* d_iq,
                                     void* cuda_stream) {
    if (!d || !d_iq || n_streams <= 0 || n <= 0) return NVX_ERR_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return NVX_ERR_CUDA;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const long long total_bits = d->bit_off[n_streams];
    std::vector<int8_t> sign((size_t)total_bits);
    std::vector<int> cum((size_t)total_bits);
    std::vector<long long> start((size_t)n_streams);
    for (int s = 0; s < n_streams; ++s) {
        int acc = 0;
        for (long long b = d->bit_off[s]; b < d->bit_off[s + 1]; ++b) {
            const int sg = d->bits[b] ? -1 : 1;
            sign[(size_t)b] = (int8_t)sg;
            cum[(size_t)b] = acc;
            acc += sg;
        }
        start[(size_t)s] = llround((double)d->start_s[s] * NVX_FS_HZ);
    }
    SynthArgs a;
    int8_t* d_sign = nullptr; int* d_cum = nullptr; long long* d_off = nullptr; long long* d_start = nullptr;
    float *d_offhz = nullptr, *d_amp = nullptr, *d_sigma = nullptr;
    cudaError_t e = cudaSuccess;
    auto up = [&](void** dst, const void* src, size_t bytes) {
        if (e != cudaSuccess) return;
        e = cudaMalloc(dst, bytes ? bytes : 1);
        if (e == cudaSuccess && bytes) e = cudaMemcpyAsync(*dst, src, bytes, cudaMemcpyHostToDevice, st);
    };
    up((void**)&d_sign, sign.data(), sign.size());
    up((void**)&d_cum, cum.data(), cum.size() * sizeof(int));
    up((void**)&d_off, d->bit_off, sizeof(long long) * (size_t)(n_streams + 1));
    up((void**)&d_start, start.data(), sizeof(long long) * (size_t)n_streams);
    up((void**)&d_offhz, d->offset_hz, sizeof(float) * (size_t)n_streams);
    up((void**)&d_amp, d->amplitude, sizeof(float) * (size_t)n_streams);
    up((void**)&d_sigma, d->noise_sigma, sizeof(float) * (size_t)n_streams);
    if (e == cudaSuccess) {
        a.sign = d_sign; a.cum = d_cum; a.bit_off = d_off; a.offset_hz = d_offhz; a.start = d_start;
        a.amplitude = d_amp; a.sigma = d_sigma; a.seed = d->seed; a.t0 = t0; a.n = n; a.streams = n_streams;
        a.out = static_cast<float2*>(d_iq);
        long long bx = (n + 255) / 256;
        if (bx > 2048) bx = 2048;
        synth_kernel<<<dim3((unsigned)bx, (unsigned)n_streams), 256, 0, st>>>(a);
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    }
    cudaFree(d_sign); cudaFree(d_cum); cudaFree(d_off); cudaFree(d_start); cudaFree(d_offhz); cudaFree(d_amp); cudaFree(d_sigma);
    return e == cudaSuccess ? NVX_OK : NVX_ERR_CUDA;
}
