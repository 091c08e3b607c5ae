// kernels.h — internal launcher prototypes (one per kernel family).
#pragma once
#include "common.h"

namespace tb200 {

int launch_decimate_exact(tb200_handle h, const void *x, int dtype, int n_chunks, long n, long chunk_stride,
                          int q, const DecimDesign &d, void *scratch, void *dec_out, long n_dec, cudaStream_t st);
int launch_channel_exact(tb200_handle h, const void *x, int dtype, int n_chunks, long n, long in_stride,
                         const double *freq_off, double fs, const ChanDesign &cd, int filtered_flag,
                         double *scratch, double *filtered, long out_stride, cudaStream_t st);
int launch_symbols(tb200_handle h, const double *filtered, int n_chunks, long n, long in_stride, int sps,
                   int phase_step, double *symbols, uint8_t *dibits, long out_stride, int32_t *n_dibits,
                   int32_t *n_symbols, int32_t *phase, cudaStream_t st, const double *phase_partial = nullptr,
                   const double *phase_tail = nullptr, int n_tiles = 0);
int launch_slicer(tb200_handle h, const double *symbols, long n, uint8_t *dibits, cudaStream_t st);

int launch_symbols_to_bits(tb200_handle h, const uint8_t *sym, int n_streams, long stride, const int32_t *n_sym,
                           int n_fixed, uint8_t *bits, uint8_t *mapped, long bits_stride, cudaStream_t st);
int launch_sync_corr(tb200_handle h, const uint8_t *bits, int n_streams, long bits_stride, const int32_t *n_sym,
                     int n_bits_fixed, uint8_t *m1, uint8_t *m2, cudaStream_t st);
int launch_find_sync(tb200_handle h, const uint8_t *m1, const uint8_t *m2, int n_bits, double thr, int32_t *pos,
                     int max_pos, int32_t *count, int32_t *max_match, cudaStream_t st);
int launch_decode(tb200_handle h, const uint8_t *bits, const uint8_t *mapped, const uint8_t *m1, const uint8_t *m2,
                  int n_streams, long bits_stride, const int32_t *n_sym, tb200_frame *frames, int max_frames,
                  int32_t *n_frames, int64_t *stats, cudaStream_t st);
int launch_tail_fused(tb200_handle h, const double *filtered, int n_chunks, long n, long in_stride, int sps, int phase_step,
                      double *symbols, uint8_t *dibits, long out_stride, int32_t *n_dibits, int32_t *phase,
                      const double *phase_partial, const double *phase_tail, int n_tiles, tb200_frame *frames, int max_frames,
                      int32_t *n_frames, int64_t *stats, cudaStream_t st);
int launch_decode_fused(tb200_handle h, const uint8_t *dibits, int n_streams, long stride, const int32_t *n_sym,
                        tb200_frame *frames, int max_frames, int32_t *n_frames, int64_t *stats, cudaStream_t st);
int launch_burst_batch(tb200_handle h, const uint8_t *dibits, int64_t n, uint8_t *rec, uint8_t *full,
                       int64_t *stats, cudaStream_t st);
int launch_crc16_rows(tb200_handle h, const uint8_t *bits, int n_rows, int n_bits, int row_stride, uint16_t *crc,
                      cudaStream_t st);
int launch_psd(tb200_handle h, const float *iq, int64_t n_frames, int64_t frame_stride, int n_fft, double fs,
               double alpha, int first, float *power, float *smoothed, double *ema_state, double *stats,
               double *scratch_db, cudaStream_t st, int use_f32 = 0, int avg = 1);
int launch_etsi_aach(tb200_handle h, const int8_t *soft, long n, const uint32_t *gen_rows, uint32_t scramb_init, uint16_t *bits,
                     int32_t *metric, int32_t *margin, cudaStream_t st);
int launch_etsi_decode_custom(tb200_handle h, const int8_t *soft, long n, int K1, int K2, int K3, int a, int rate, int crc,
                              uint32_t scramb_init, uint8_t *bits, uint8_t *crc_ok, int32_t *metric, cudaStream_t st);
int launch_etsi_decode(tb200_handle h, const int8_t *soft, long n, int channel, uint32_t scramb_init, uint8_t *bits,
                       uint8_t *crc_ok, int32_t *metric, cudaStream_t st);
int launch_etsi_demod(tb200_handle h, const float2 *iq, int n_streams, long n, long stride, double fs, float *taps_dev,
                      float2 *sym, float *soft, long soft_stride, int32_t *n_soft, void *scratch, size_t scratch_bytes, cudaStream_t st);
int64_t etsi_demod_scratch_bytes(int n_streams, long n, double fs, int want_symbols);
int launch_etsi_find_bursts(tb200_handle h, const float *soft, int n_streams, long soft_stride, const int32_t *n_soft,
                            float thr, float *metric, int8_t *blocks, int32_t *burst_pos, int32_t *n_bursts, int max_bursts,
                            cudaStream_t st);
int launch_etsi_find_sequence(tb200_handle h, const float *soft, int n_streams, long soft_stride, const int32_t *n_soft, int seq,
                              float thr, int gap, float *metric, int32_t *pos, float *corr, int32_t *n_hits, int max_hits,
                              cudaStream_t st);
int launch_etsi_extract(tb200_handle h, const float *soft, int n_streams, long soft_stride, const int32_t *n_soft,
                        const int32_t *burst_pos, const int32_t *n_bursts, int max_bursts, int off_a, int len_a, int off_b, int len_b,
                        int8_t *out, cudaStream_t st);
int launch_etsi_afc(tb200_handle h, float *soft, int n_streams, long soft_stride, const int32_t *n_soft, float *theta, cudaStream_t st);
int launch_etsi_find_sync(tb200_handle h, const float *soft, int n_streams, long soft_stride, const int32_t *n_soft, float thr,
                          float *metric, int8_t *sb1, int8_t *bkn2, int32_t *burst_pos, int32_t *n_bursts, int max_bursts,
                          cudaStream_t st);
int launch_channelizer(tb200_handle h, const float2 *x, long n, int M, float *proto_dev, float2 *tw_dev, float2 *out,
                       long out_stride, long n_frames, cudaStream_t st);
int launch_polyphase_bank(tb200_handle h, const float2 *x, long n, int M, int N1, int N2, int P, int l0, int hop, long n_frames,
                          const float *proto, const float2 *tw, float2 *out, long out_stride, cudaStream_t st);
int launch_refchan(tb200_handle h, const float2 *x, long n, int M, int q, double fs, const DecimDesign &d, float2 *out,
                   long out_stride, cudaStream_t st);
int launch_sc16_to_c64(tb200_handle h, const void *in, long n, void *out, cudaStream_t st);
int launch_resample(tb200_handle h, const void *x, int dtype, long n, long num, void *y, double *scratch, size_t scratch_bytes,
                    cudaStream_t st);
int launch_freq_shift(tb200_handle h, const void *x, int dtype, long n, double f, double fs, double *out, cudaStream_t st);
bool chan_fast_supported(double bandwidth, double fs);
int launch_detect(tb200_handle h, const void *x, int dtype, int n_dwells, long n, long stride, double fs, int num_windows,
                  tb200_detect_stats *out, void *scratch, size_t scratch_bytes, cudaStream_t st);
int launch_voice_blocks(tb200_handle h, int mode, const void *src, long src_stride, const int32_t *n_src,
                        const int32_t *frame_stream, const int32_t *frame_pos, long n_frames, float scale,
                        int16_t *blocks, uint8_t *valid, cudaStream_t st);

}  // namespace tb200
