// Deterministic synthetic ERA5-like field generated on the device: any shard
// (t0, level) of the month-long benchmark field reproduces the same values without ever
// materialising the 110 GB array on the host.  Counter-based noise keyed by the global voxel
// index (splitmix64 -> Box-Muller), so the result does not depend on the launch geometry.
#include "common.cuh"

namespace {

__device__ __forceinline__ unsigned long long splitmix64(unsigned long long z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

__global__ void synth_kernel(float* __restrict__ x, int T, int H, int W, int t0, int level,
                             unsigned long long seed) {
  const long long n = (long long)T * H * W;
  const double PI = 3.14159265358979323846;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const int w = (int)(i % W);
    const int h = (int)((i / W) % H);
    const int t = (int)(i / ((long long)W * H));
    const double lat = PI / 2 - h * (PI / (H - 1));
    const double lon = w * (2 * PI / W) - PI;
    const double tt = (double)(t + t0);
    double v = 288.0 - 45.0 * sin(lat) * sin(lat) + 6.0 * sin(3 * lon + 0.4 * lat) * cos(lat) +
               3.0 * cos(2 * PI * tt / 24.0 - lon) * cos(lat) + 2.0 * sin(11 * lon + 7 * lat) +
               0.8 * sin(37 * lon - 29 * lat + 0.3 * tt) - 6.5 * level;
    const unsigned long long gidx = ((unsigned long long)(t + t0) * H + h) * W + w;
    const unsigned long long r = splitmix64(seed ^ splitmix64(gidx + 0x1000003ull * (unsigned long long)level));
    const double u1 = ((r >> 11) + 1.0) * (1.0 / 9007199254740993.0);
    const double u2 = (splitmix64(r) >> 11) * (1.0 / 9007199254740992.0);
    v += 0.15 * sqrt(-2.0 * log(u1)) * cos(2 * PI * u2);
    x[i * 2] = (float)v;
    x[i * 2 + 1] = (sin(2 * lon + 1) + cos(3 * lat) > 0.3) ? 1.f : 0.f;
  }
}

}  // namespace

extern "C" int lc_synth_field(float* d_x, int T, int H, int W, int t0, int level, uint64_t seed,
                              void* stream) {
  LC_REQUIRE(d_x && T > 0 && H > 1 && W > 0, "bad argument");
  synth_kernel<<<148 * 8, 256, 0, as_stream(stream)>>>(d_x, T, H, W, t0, level, seed);
  LC_LAUNCH_CHECK();
  return LC_OK;
}
