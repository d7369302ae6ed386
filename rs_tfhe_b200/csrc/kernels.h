// kernels.h -- internal launch interface between engine.cu and the kernels.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/tfhe_b200.h"
#include "br_core.cuh"

enum { BR_OUT_TRLWE = 0, BR_OUT_EXTRACT = 1, BR_OUT_EXTRACT2 = 2 };

// K0+K3 (blind_rotate.cu)
struct BrArgs {
  const cplx *bsk;          // device order: cplx[n][2l][8][2][64] (br_core.cuh)
  const cplx *bsk2;         // same rows, thread slots permuted for the TMEM-exchange kernel (or NULL)
  const cplx *bsk3;         // same rows in the 128-thread kernel's order [kd][o][T] (brs_core.cuh), or NULL
  const cplx *tw_s;         // [128][20] per-thread constants of the 128-thread kernel
  const cplx *tw_a;         // [64][8]
  const cplx *tw_b;         // [8][8]
  const uint32_t *tv;       // test vectors u32[slot][2][N]; slot 0 = cloud-key test vector
  const int32_t *tv_index;  // per-ciphertext slot, or NULL -> tv_default
  int tv_default;
  const uint32_t *in;       // op>=0 or ops: [count][2][n+1]; else [count][n+1]
  const uint8_t *ops;       // per-ciphertext gate or NULL
  int op;                   // gate for all, or -1: plain bootstrap of `in`
  uint32_t *out;            // per out_mode: [count][2][N] | [count][N+1] | [count][n+1]
  int out_mode;
  uint32_t n;
  uint32_t offset;          // CloudKey.decomposition_offset, used verbatim
  size_t count;
};
bool br_supported(uint32_t l, uint32_t bgbit);
bool br_uses_permuted_key();  // the selected throughput kernel reads BrArgs::bsk2
bool br_uses_s_key();         // the selected throughput kernel reads BrArgs::bsk3 / tw_s
cudaError_t br_launch(uint32_t l, uint32_t bgbit, const BrArgs &args, int num_sms,
                      cudaStream_t stream, int *launched = nullptr);
// one ciphertext per CTA over l groups of 128 threads (blind_rotate_s.cu): small batches, dependent chains
cudaError_t br_launch_latency_s(uint32_t l, uint32_t bgbit, const BrArgs &args, int num_sms,
                                cudaStream_t stream);
// one ciphertext per 2-CTA thread-block cluster, partial spectra exchanged through distributed shared memory
cudaError_t br_launch_cluster(uint32_t l, uint32_t bgbit, const BrArgs &args, int num_sms,
                              cudaStream_t stream);
// 128-thread-per-ciphertext throughput kernel (blind_rotate_s.cu)
cudaError_t br_launch_s(uint32_t l, uint32_t bgbit, const BrArgs &args, int num_sms,
                        cudaStream_t stream);

// FFTProcessor seam (fft_seam.cu): standalone batch transforms on the blind rotation's passes
enum { MODE_IFFT = 0, MODE_FFT = 1, MODE_POLYMUL = 2 };
cudaError_t fft_seam_launch(int mode, const cplx *tw_s, const void *in_a, const uint32_t *in_b, void *out,
                            size_t count, int num_sms, cudaStream_t stream);

// K4 (keyswitch.cu)
struct KsArgs {
  const uint32_t *ksk;   // device order: u32[rows+1][stride]; last row = zeros
  const uint32_t *ext;   // [count][N+1] extracted level-1 samples
  uint32_t *out;         // [count][n+1]
  uint32_t n, basebit, iks_t, stride, zero_row;
  uint32_t n_in;         // input dimension: ext is [count][n_in+1] (N for K4, n for re-encryption)
  size_t count;
};
cudaError_t ks_launch(const KsArgs &args, cudaStream_t stream);
// a handful of ciphertexts: the (i, j) terms of each split over ~2 CTAs per SM, partial sums combined with atomics
cudaError_t ks_small_launch(const KsArgs &args, int num_sms, cudaStream_t stream);
static inline uint32_t ks_stride(uint32_t n) { return (n + 1 + 3) & ~3u; }

// K4 on the tcgen05 tensor cores (keyswitch_umma.cu): basebit 2..6 (gate sets, UINT1-6)
struct KsUmmaArgs {
  const uint8_t *key;   // operand tiles, ku_layout.h: [n tile][stage][kstep 2][half 2][240 x 32 B canonical]
  const uint32_t *ext;  // [count][N+1]
  uint32_t *out;        // [count][n+1]
  uint32_t n, iks_t, basebit;
  size_t count;
};
bool ks_umma_supported(uint32_t basebit, uint32_t t);
size_t ks_umma_key_bytes(uint32_t n, uint32_t t, uint32_t basebit);
cudaError_t ks_umma_launch(const KsUmmaArgs &args, cudaStream_t stream);
cudaError_t ksk_umma_relayout_launch(const uint32_t *blob_rows, uint32_t stride, uint8_t *dst,
                                     uint32_t n, uint32_t t, uint32_t basebit, cudaStream_t stream);

// layout / small kernels (aux.cu)
cudaError_t bsk_relayout_launch(const double *src_ref, cplx *dst, uint32_t n, uint32_t l2,
                                cudaStream_t stream);
cudaError_t bsk_permute_launch(const cplx *src, cplx *dst, size_t rows, cudaStream_t stream);
cudaError_t bsk_permute_s_launch(const cplx *src, cplx *dst, size_t rows, cudaStream_t stream);
cudaError_t ksk_relayout_launch(const uint32_t *src_ref, uint32_t *dst, uint32_t rows, uint32_t n,
                                uint32_t stride, cudaStream_t stream);
cudaError_t lut_generate_launch(const uint32_t *d_f_table, uint32_t modulus, double scale,
                                uint32_t *d_tv_slot, cudaStream_t stream);
cudaError_t extract_launch(const uint32_t *d_trlwe, uint32_t *d_ext, size_t count,
                           cudaStream_t stream);
cudaError_t fp64_probe_launch(double *d_sink, int blocks, int iters, bool three_operands, cudaStream_t stream);

// K6 (keygen.cu): cloud-key generation on the device
cudaError_t keygen_launch(const cplx *tw_a, const cplx *tw_b, const uint32_t *d_s0,
                          const uint32_t *d_s1, cplx *d_s1_spec, cplx *d_bsk, uint32_t *d_ksk_ref,
                          uint32_t n, uint32_t l, uint32_t bgbit, uint32_t basebit, uint32_t t,
                          double alpha_lv0, double alpha_lv1, const uint32_t key256[8], cudaStream_t stream);
