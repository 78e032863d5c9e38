// |0...0> fill and export of the uint8 tableau store.
// Part of libsdimb (sdim_b200/csrc); included by sdimb.cu inside its anonymous namespace.
#pragma once

__global__ void init_kernel(uint8_t* tab, int n, int np, int W, int64_t row_bytes, int64_t shot_bytes, int64_t shots) {
  for (int64_t shot = blockIdx.x; shot < shots; shot += gridDim.x) {
    uint8_t* T = tab + shot * shot_bytes;
    uint4* v = reinterpret_cast<uint4*>(T);
    for (int64_t i = threadIdx.x; i < shot_bytes / 16; i += blockDim.x) v[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    for (int q = threadIdx.x; q < n; q += blockDim.x) {
      T[(int64_t)q * row_bytes + W + q] = 1;
      T[(int64_t)q * row_bytes + np + q] = 1;
    }
    __syncthreads();
  }
}

__global__ void export_kernel(const uint8_t* T, int n, int np, int W, int64_t row_bytes, int64_t phase_off,
                              int64_t* x, int64_t* z, int64_t* ph, int64_t* dx, int64_t* dz, int64_t* dph) {
  const int64_t total = (int64_t)n * n;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int q = (int)(i / n), g = (int)(i % n);
    const uint8_t* row = T + (int64_t)q * row_bytes;
    x[i] = row[g];
    z[i] = row[W + g];
    dx[i] = row[np + g];
    dz[i] = row[W + np + g];
    if (q == 0) {
      ph[g] = T[phase_off + g];
      dph[g] = T[phase_off + np + g];
    }
  }
}
