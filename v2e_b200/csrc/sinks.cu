// Event-sink row conversions on the device: the packed float32 rows [t, x, y, p] that the
// pixel model emits become what the reference's writers put into files, so that the host copies 8 (AEDAT-2.0)
// or 16 (HDF5) bytes per event straight into the sink instead of converting row by row in numpy.
//
// Replaces (reference = SensorsINI/v2e, /root/reference):
//   v2ecore/emulator.py:953-959            HDF5 "events" dataset rows: uint32 [t_us, x, y, p01]
//   v2ecore/output/aedat2_output.py:133-165  AEDAT-2.0: int32 address / int32 timestamp pairs, big endian
// File headers, h5py and the '#'-first-byte check of aedat2_output.py:166-172 stay with the caller.
// Both conversions are pure streaming (16 B read per event): HBM bound, one thread per event.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/v2e_b200.h"

extern int v2e_set_error(int code, const char *fmt, const char *detail);

namespace {

// numpy float32 -> uint32 / int32 casts truncate toward zero
__device__ __forceinline__ uint32_t f2u_trunc(float v) { return (uint32_t)(int64_t)v; }

__global__ void __launch_bounds__(256) h5_rows_kernel(const float4 *__restrict__ ev, uint64_t n, uint4 *__restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 e = ev[i];
    // emulator.py:955-958: temp[:,0] *= 1e6 in float32; p == -1 -> 0; astype(uint32)
    const float t_us = __fmul_rn(e.x, 1e6f);
    const float p = e.w == -1.0f ? 0.0f : e.w;
    out[i] = make_uint4(f2u_trunc(t_us), f2u_trunc(e.y), f2u_trunc(e.z), f2u_trunc(p));
}

__global__ void __launch_bounds__(256)
aedat2_kernel(const float4 *__restrict__ ev, uint64_t n, int size_x, int size_y, int xs, int ys, int ps, int flip_x,
              int flip_y, uint2 *__restrict__ out, unsigned long long *__restrict__ n_on) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int on = 0;
    if (i < n) {
        const float4 e = ev[i];
        const int32_t t = (int32_t)__fmul_rn(1e6f, e.x);              // aedat2_output.py:144
        int32_t x = (int32_t)e.y, y = (int32_t)e.z;
        if (flip_x) x = (size_x - 1) - x;                             // :148
        if (flip_y) y = (size_y - 1) - y;                             // :150
        const int32_t p = (int32_t)__fdiv_rn(__fadd_rn(e.w, 1.0f), 2.0f);   // :151
        const uint32_t a = ((uint32_t)x << xs) | ((uint32_t)y << ys) | ((uint32_t)p << ps);   // :153
        // address then timestamp, each byte-swapped to big endian (:160-162)
        out[i] = make_uint2(__byte_perm(a, 0, 0x0123), __byte_perm((uint32_t)t, 0, 0x0123));
        on = p != 0;
    }
    if (n_on) {
        const unsigned m = __ballot_sync(0xffffffffu, on);
        if ((threadIdx.x & 31) == 0 && m) atomicAdd(n_on, (unsigned long long)__popc(m));
    }
}

}  // namespace

extern "C" int v2e_events_to_h5_rows(const float *events_dev, uint64_t n, uint32_t *rows_dev, void *stream) {
    if (n == 0) return V2E_OK;
    if (!events_dev || !rows_dev) return v2e_set_error(V2E_E_INVALID, "null argument%s", "");
    if (((uintptr_t)events_dev | (uintptr_t)rows_dev) & 15) return v2e_set_error(V2E_E_INVALID, "buffers must be 16-byte aligned%s", "");
    const uint64_t blocks = (n + 255) / 256;
    if (blocks > 0x7fffffffull) return v2e_set_error(V2E_E_INVALID, "too many events for one call%s", "");
    h5_rows_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const float4 *)events_dev, n, (uint4 *)rows_dev);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return v2e_set_error(V2E_E_CUDA, "h5_rows_kernel: %s", cudaGetErrorString(e));
    return V2E_OK;
}

extern "C" int v2e_events_to_aedat2(const float *events_dev, uint64_t n, int size_x, int size_y, int x_shift,
                                    int y_shift, int pol_shift, int flip_x, int flip_y, uint32_t *words_dev,
                                    uint64_t *n_on_dev, void *stream) {
    if (n == 0) return V2E_OK;
    if (!events_dev || !words_dev) return v2e_set_error(V2E_E_INVALID, "null argument%s", "");
    if (((uintptr_t)events_dev & 15) || ((uintptr_t)words_dev & 7)) return v2e_set_error(V2E_E_INVALID, "misaligned buffer%s", "");
    if (x_shift < 0 || y_shift < 0 || pol_shift < 0 || x_shift > 31 || y_shift > 31 || pol_shift > 31)
        return v2e_set_error(V2E_E_INVALID, "bad shift%s", "");
    const uint64_t blocks = (n + 255) / 256;
    if (blocks > 0x7fffffffull) return v2e_set_error(V2E_E_INVALID, "too many events for one call%s", "");
    aedat2_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const float4 *)events_dev, n, size_x, size_y, x_shift,
                                                                       y_shift, pol_shift, flip_x, flip_y, (uint2 *)words_dev,
                                                                       (unsigned long long *)n_on_dev);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return v2e_set_error(V2E_E_CUDA, "aedat2_kernel: %s", cudaGetErrorString(e));
    return V2E_OK;
}
