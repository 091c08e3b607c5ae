/*
 * tetra_b200.h — C-ABI of libtetra_b200.so, the B200 (sm_100a) implementation of TetraEar's
 * IQ -> bursts receive path.
 *
 * The reference (WizzardDr/TetraEar-BladeRF) is pure Python and has no FFI of its own; the
 * boundary this library replaces is the Python class API its callers import
 * (tetraear/ui/modern.py:194-195, tetraear/signal/scanner.py:16-17).  Each entry point below
 * names the reference method(s) whose arithmetic it performs.  The ctypes binding a
 * maintainer adds on the reference side is shown in INTEGRATION.md.
 *
 * Conventions
 *   - every function returns 0 on success, < 0 on error (TB200_E_*); the message of the last
 *     error on the calling thread is tb200_last_error().  Nothing throws or aborts.
 *   - pointers named d_* are DEVICE pointers owned by the caller (e.g. torch tensors'
 *     data_ptr()); pointers named h_* are HOST pointers.  `stream` is a cudaStream_t passed
 *     as void* (NULL = legacy default stream).  Calls are asynchronous on `stream` unless the
 *     name ends in _host (those synchronise `stream` before returning).
 *   - all scratch memory is allocated by tb200_create for the declared capacity; no entry
 *     point allocates afterwards.  One handle per thread of execution; distinct handles are
 *     independent (the reference builds fresh instances per thread, scanner.py:164-171).
 *   - complex samples are interleaved (re, im); complex64 = 2 x float, complex128 = 2 x double.
 */
#ifndef TETRA_B200_H
#define TETRA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TB200_OK 0
#define TB200_E_ARG (-1)      /* bad argument / exceeds handle capacity */
#define TB200_E_CUDA (-2)     /* CUDA runtime error (message has the detail) */
#define TB200_E_NODEV (-3)    /* no CUDA device: there is no CPU fallback */
#define TB200_E_ALLOC (-4)

#define TB200_DTYPE_C64 0
#define TB200_DTYPE_C128 1
/* interleaved int16 (I, Q) as the bladeRF delivers it (SC16Q11); converted on the device exactly like
 * BladeRFCapture.read_samples does on the host: complex64(I, Q) / 32768  (signal/capture.py:255-269) */
#define TB200_DTYPE_SC16 2

/* decimator arithmetic: EXACT = scipy's float32 op order, sequential per (chunk, I/Q) lane,
 * bit-identical to the reference; FAST = time-tiled, same filter, float32/float64 mixed,
 * soft symbols within 1e-4 of full scale */
#define TB200_MODE_EXACT 0
#define TB200_MODE_FAST 1

#define TB200_MAX_PHASES 32
#define TB200_FRAME_BITS 510
#define TB200_BURST_SYMS 255

typedef struct tb200_handle_s *tb200_handle;

typedef struct tb200_config {
    double sample_rate;   /* SignalProcessor(sample_rate), processor.py:21-33 */
    int32_t max_chunks;   /* capacity: chunks per tb200_process_batch call */
    int32_t max_chunk_len;/* capacity: samples per chunk */
    int32_t mode;         /* TB200_MODE_* */
    int32_t device;       /* CUDA device ordinal */
    int32_t max_fft;      /* capacity of the PSD path (0 -> 4096) */
    int32_t max_bursts;   /* capacity of tb200_burst_batch_host staging (0 -> none) */
    int32_t max_stream_len; /* capacity: dibits per stream of the decoder entry points (0 -> derived) */
    int32_t max_channels; /* capacity of tb200_channelize (0 -> no channelizer workspace) */
} tb200_config;

/* What process() will do for (sample_rate, chunk_len): mirrors the control flow of
 * SignalProcessor.process, processor.py:239-273 (decimate only if fs > 480 kHz and
 * len > 27; filter only if len > 15; sps = int(rate/18000); phase step = max(1, sps/8)). */
typedef struct tb200_plan {
    int32_t q;            /* decimation factor actually applied (1 = none) */
    int32_t n_dec;        /* samples after decimation */
    int32_t filtered;     /* 1 if the Butterworth filtfilt runs */
    int32_t sps;          /* int(rate / 18000) */
    int32_t phase_step;   /* max(1, sps / 8) */
    int32_t max_symbols;  /* upper bound of symbols per chunk (phase 0) */
    double rate;          /* sample rate after decimation */
} tb200_plan;

/* one frame candidate of TetraDecoder.decode (core/decoder.py:861-888 + decode_frame :890-992) with the
 * fixed-offset MAC header fields of parse_mac_pdu and the payload statistics its callers test — 156 bytes */
typedef struct tb200_frame {
    int32_t sync_pos;     /* bit position of the training sequence */
    int32_t start_pos;    /* sync_pos - 216 */
    int32_t number;       /* start_pos / 510 (frame['number']); timeslot = number % 4 */
    uint8_t type;         /* bits[0:2]  (frame['type'], decoder.py:906) */
    uint8_t enc_mode;     /* bits[2:4]  (decoder.py:909) */
    uint8_t burst_type;   /* BurstType value: 2 NormalDownlink, 5 Synchronization (protocol.py:34-41) */
    uint8_t crc_ok;       /* TetraBurst.crc_ok, protocol.py:292-329 */
    uint8_t mac_ok;       /* parse_mac_pdu() would return a PDU (protocol.py:349-596) */
    uint8_t kept;         /* 1 unless dropped by decoder.py:1092-1100 (mac None and CRC bad) */
    uint8_t mac_fill;     /* MAC header fill-bit indicator, data bit 4 (protocol.py:402, :454, :510) */
    uint8_t mac_length;   /* length field: data bits 29..34 (MAC-RESOURCE, :418-421) or 5..10 (MAC-END, :514-517), else 0 */
    uint32_t mac_address; /* 24-bit address, data bits 5..28, MAC-RESOURCE only (:410-413), else 0 */
    uint8_t mac_data_bytes; /* len(MacPDU.data): the payload bits of protocol.py:423-443 / :456-462 / :496-501 / :519-537 packed to bytes */
    uint8_t mac_unique;   /* len(set(MacPDU.data)): numerator of the entropy heuristic, decoder.py:1036-1053 */
    uint8_t mac_b0, mac_b1; /* MacPDU.data[0], [1] (0 when absent): the SDS protocol-id pre-filters, protocol.py:820-900 */
    uint8_t bits[64];     /* the 510 frame bits (frame['bits']), MSB-first packed */
    uint8_t data[64];     /* TetraBurst.data_bits packed: 216 (normal) or 510 (sync) bits */
} tb200_frame;

/* ---- library ------------------------------------------------------------------------- */
const char *tb200_last_error(void);
int tb200_version(void);
int tb200_device_count(void);   /* 0 when no usable CUDA device */

/* ---- filter design: host only, no device needed -------------------------------------- */
/* scipy.signal.decimate's filter for factor q as the reference gets it (processor.py:254):
 * cheby1(8, 0.05, 0.8/q) SOS rows [b0 b1 b2 1 a1 a2] and sosfilt_zi, in float64 and in the
 * float32 (complex64 arithmetic) form used for complex64 IQ. */
int tb200_design_decimator(int q, double sos64[24], double zi64[8], float sos32[24], float zi32[8]);
/* butter(4, clip((bw/2)/(fs/2), .01, .99)) and lfilter_zi, processor.py:70-79 */
int tb200_design_channel_filter(double bandwidth, double fs, double b[5], double a[5], double zi[4]);
int tb200_plan_process(double sample_rate, int chunk_len, tb200_plan *plan);

/* ---- handle -------------------------------------------------------------------------- */
int tb200_create(const tb200_config *cfg, tb200_handle *out);
int tb200_destroy(tb200_handle h);
int tb200_set_sample_rate(tb200_handle h, double sample_rate); /* attr is mutated live, modern.py:1857 */
/* options: TB200_OPT_CHAN_F32 = 1: in FAST mode evaluate the Butterworth channel filter's recurrences in
 * float32 instead of float64 (soft symbols then agree with the reference to ~1e-6 of full scale instead
 * of ~1e-13 for that stage; the 1e-4 budget of FAST mode is dominated by the reference's own float32
 * decimator either way; the NCO of a lane's 8 samples is then one float64 sincos + a float32 rotation).  Default 0. */
#define TB200_OPT_CHAN_F32 1
/* TB200_OPT_PSD_F32 = 2: PSD frames with a float32 FFT and float32 dB conversion (amplitude error ~1e-7 of the
 * frame peak, i.e. far inside the 1e-4-of-full-scale budget; dB values of bins > 100 dB below the frame peak lose
 * digits).  Default 0 = float64 like numpy. */
#define TB200_OPT_PSD_F32 2
/* TB200_OPT_DECIM_TC = 3: in FAST mode run the decimator's linear part (q-tap modal FIRs and the in-lane recurrences)
 * as split-TF32 tcgen05 MMAs (dsp_tc.cu) when q is 8 or 10, the input is complex64, 16-byte aligned with an even chunk
 * stride and a chunk holds at least 1536 outputs; other geometries use the FFMA2 kernel of dsp_fast.cu.  Same
 * filter, same float32-level accuracy (<= 1e-4 of full scale vs the reference; 7e-7 of full scale from the exact-arithmetic
 * filter).  Default 0: on this round's B200 measurements the FFMA2 kernel is the faster of the two. */
#define TB200_OPT_DECIM_TC 3
/* TB200_OPT_PSD_AVG = 4: linear (Welch) averaging for tb200_psd_batch: value K >= 1; the n_frames input frames are taken
 * K at a time and every OUTPUT frame (floor(n_frames / K) of them; power, smoothed and stats rows) is
 * 20 log10( sqrt(mean_k (|X_k| / N)^2) + 1e-20 ), i.e. the reference's dBFS formula (ui/modern.py:1941) on the RMS of the K
 * amplitude spectra; K = 1 (default) is the reference's per-frame spectrum.  The dB-domain EMA, if requested, runs over
 * the averaged frames. */
#define TB200_OPT_PSD_AVG 4
/* TB200_OPT_ETSI_OPEN_LOOP = 5: tb200_etsi_demod_batch with both timing-loop gains zero: symbol k is taken at
 * t = 16 + (k + 1) fs / 18000 samples exactly, so its matched-filter output can be checked against numpy.convolve with
 * the same taps (tb200_debug_read region 4 returns the 32 x 32 interpolating RRC tap table).  Default 0. */
#define TB200_OPT_ETSI_OPEN_LOOP 5
/* TB200_OPT_ETSI_SEG_LANES = 6: form of tb200_etsi_demod_segmented.  0 (default): two stages — the RRC matched filter once
 * per input sample for all streams, then one thread per (stream, segment) running the timing loop on 8-tap interpolations
 * of its output; 8: eight lanes per segment with the 32-tap interpolating matched filter inside the loop. */
#define TB200_OPT_ETSI_SEG_LANES 6
int tb200_set_option(tb200_handle h, int option, int value);
size_t tb200_workspace_bytes(tb200_handle h);
/* number of kernels this handle has launched so far (bench.py's gpu_launches) */
int64_t tb200_launch_count(tb200_handle h);

/* ---- SignalProcessor.process over a batch of equal-length chunks (processor.py:221-273) -
 * d_iq        [n_chunks][chunk_stride] complex samples (dtype: TB200_DTYPE_*)
 * d_freq_off  [n_chunks] double, Hz, or NULL for 0 (freq_offset; NCO skipped where == 0)
 * d_dibits    [n_chunks][out_stride] uint8   demodulated symbols 0..3
 * d_symbols   [n_chunks][out_stride] complex128 (self.symbols) or NULL
 * d_n_dibits  [n_chunks] int32  dibits produced per chunk (symbols = dibits + 1 when > 0)
 * d_phase     [n_chunks] int32  timing phase picked (may be NULL)                          */
int tb200_process_batch(tb200_handle h, const void *d_iq, int dtype, int n_chunks, int chunk_len,
                        int64_t chunk_stride, const double *d_freq_off, uint8_t *d_dibits,
                        void *d_symbols, int out_stride, int32_t *d_n_dibits, int32_t *d_phase,
                        void *stream);
/* process() followed by TetraDecoder.decode() of every chunk (decoder.py:835-888) — what CaptureThread.run does per
 * chunk (ui/modern.py:2032-2044) — as one call: the outputs of tb200_process_batch plus those of tb200_decode_batch
 * (d_frames [n_chunks][max_frames], d_n_frames [n_chunks], d_stats [3] or NULL).  Where a chunk's dibits fit the
 * one-launch decoder (out_stride <= 6000, max_frames <= 64) the timing pick, slicer, sync search and frame records are
 * ONE kernel launch per batch; otherwise the two calls run back to back.  Same results either way. */
int tb200_process_decode_batch(tb200_handle h, const void *d_iq, int dtype, int n_chunks, int chunk_len,
                               int64_t chunk_stride, const double *d_freq_off, uint8_t *d_dibits,
                               void *d_symbols, int out_stride, int32_t *d_n_dibits, int32_t *d_phase,
                               tb200_frame *d_frames, int max_frames, int32_t *d_n_frames, int64_t *d_stats,
                               void *stream);
/* same, HOST buffers in and out: H2D of the IQ, the kernels, D2H of dibits / counts / symbols,
 * all on `stream`, synchronised before return. */
int tb200_process_batch_host(tb200_handle h, const void *h_iq, int dtype, int n_chunks, int chunk_len,
                             const double *h_freq_off, uint8_t *h_dibits, void *h_symbols,
                             int out_stride, int32_t *h_n_dibits, int32_t *h_phase, void *stream);

/* ---- stage-level entry points (one array each; the per-method Python API and parity tests) */
/* resample (processor.py:35-49) = scipy.signal.resample(x, num) for complex input: num complex samples of the
 * input dtype.  O(n * min(n, num)) direct DFT in float64 — not on the receive path (no caller uses it). */
int tb200_resample(tb200_handle h, const void *d_x, int dtype, int n, int num, void *d_out, void *stream);
/* scipy.signal.decimate(x, q) (processor.py:254).  out: ceil(n/q) samples of the same dtype. */
int tb200_decimate(tb200_handle h, const void *d_x, int dtype, int n, int q, void *d_out, void *stream);
/* frequency_shift (processor.py:85-100): out complex128[n] = x * exp(-j 2 pi f k / fs) */
int tb200_freq_shift(tb200_handle h, const void *d_x, int dtype, int n, double freq_offset, double fs,
                     void *d_out_c128, void *stream);
/* filter_signal (processor.py:51-83): Butterworth(4) filtfilt, complex128 out. n must be > 15. */
int tb200_filter(tb200_handle h, const void *d_x, int dtype, int n, double bandwidth, double fs,
                 void *d_out_c128, void *stream);
/* extract_symbols (processor.py:168-219): d_n_out <- count, d_phase <- best phase */
int tb200_extract_symbols(tb200_handle h, const void *d_x_c128, int n, double fs, void *d_out_c128,
                          int32_t *d_n_out, int32_t *d_phase, void *stream);
/* demodulate_dqpsk (processor.py:102-166): n complex128 symbols -> n-1 dibits */
int tb200_demod_dqpsk(tb200_handle h, const void *d_sym_c128, int n, uint8_t *d_dibits, void *stream);

/* ---- TetraDecoder (core/decoder.py) -------------------------------------------------- */
/* symbols_to_bits (:140-169): n symbols (0..3, or 0..7 -> 8-PSK map) -> 2n bits, n mapped */
int tb200_symbols_to_bits(tb200_handle h, const uint8_t *d_symbols, int n, uint8_t *d_bits,
                          uint8_t *d_mapped, void *stream);
/* find_sync (:171-295) on one bit stream: positions (<= max_pos) and count, max correlation
 * as the match COUNT out of 22 over the positions the reference visits. */
int tb200_find_sync(tb200_handle h, const uint8_t *d_bits, int n_bits, double threshold,
                    int32_t *d_pos, int max_pos, int32_t *d_count, int32_t *d_max_match, void *stream);
/* decode (:835-888) over a batch of dibit streams: threshold cascade, frame windows, header
 * bits, parse_burst, CRC, MAC validity, drop rule.
 * d_dibits [n_streams][stride]; d_n_dibits [n_streams] (device) ;
 * d_frames [n_streams][max_frames] ; d_n_frames [n_streams] = sync positions examined
 * (records with kept == 0 were dropped by the reference and must be skipped) ;
 * d_stats  int64[3] += (total_bursts, crc_pass, crc_fail)  (protocol.py:228-232) */
int tb200_decode_batch(tb200_handle h, const uint8_t *d_dibits, int n_streams, int stride,
                       const int32_t *d_n_dibits, tb200_frame *d_frames, int max_frames,
                       int32_t *d_n_frames, int64_t *d_stats, void *stream);

/* ---- TetraProtocolParser.parse_burst integer path over a batch (core/protocol.py:192-347)
 * d_dibits [n][255] ; d_rec [n][32] bytes:
 *   [0..26] first 216 data bits packed MSB-first (normal burst: bits 0..107 ++ 122..229)
 *   [27] burst_type (2 / 5)   [28] crc_ok   [29] mac_ok   [30] (type << 2) | enc_mode
 *   [31] number of 1-bits in the 22-bit window compared with the sync patterns (max match)
 * d_full [n][64] or NULL: all 510 burst bits packed (data_bits of a Synchronization burst)
 * d_stats int64[3] += (total_bursts, crc_pass, crc_fail) or NULL */
int tb200_burst_batch(tb200_handle h, const uint8_t *d_dibits, int64_t n, uint8_t *d_rec,
                      uint8_t *d_full, int64_t *d_stats, void *stream);
int tb200_burst_batch_host(tb200_handle h, const uint8_t *h_dibits, int64_t n, uint8_t *h_rec,
                           uint8_t *h_full, int64_t *h_stats, void *stream);
/* The class API's per-chunk pattern in ONE call (ui/modern.py:1919-2077: processor.process(samples), then
 * decoder.decode(demodulated)): host IQ in; dibits, soft symbols (or NULL), counts, timing phases, frame records,
 * frame counts and the three decode statistics out.  One upload, the kernels of tb200_process_decode_batch, ONE download
 * through a pinned mirror inside the handle, one synchronisation.  For the small batches of that pattern only: the
 * result region is limited to 1 MiB (n_chunks x (out_stride x 17 + max_frames x 156)), max_frames <= 64. */
int tb200_process_decode_host(tb200_handle h, const void *h_iq, int dtype, int n_chunks, int chunk_len,
                              const double *h_freq_off, uint8_t *h_dibits, void *h_symbols, int out_stride,
                              int32_t *h_n_dibits, int32_t *h_phase, tb200_frame *h_frames, int max_frames,
                              int32_t *h_n_frames, int64_t *h_stats, void *stream);
/* 1 when a call of that geometry fits the handle's result region (n chunks / streams; out_stride 0 for the decoder call) */
int tb200_host_oneshot_fits(tb200_handle h, int n, int out_stride, int max_frames, int want_symbols);
/* TetraDecoder.decode (core/decoder.py:835-888) of host dibit rows [n_streams][stride]: upload, tb200_decode_batch,
 * one download, one synchronisation; h_stats[3] = total bursts, CRC pass, CRC fail of THIS call.  Same size limits. */
int tb200_decode_batch_host(tb200_handle h, const uint8_t *h_dibits, int n_streams, int stride, const int32_t *h_n_dibits,
                            tb200_frame *h_frames, int max_frames, int32_t *h_n_frames, int64_t *h_stats, void *stream);

/* _calculate_crc16 (protocol.py:331-347) of n_rows bit rows of n_bits each -> uint16 */
int tb200_crc16_rows(tb200_handle h, const uint8_t *d_bits, int n_rows, int n_bits, int row_stride,
                     uint16_t *d_crc, void *stream);

/* ---- `etsi` mode: ETSI EN 300 392-2 clause 8 channel decoding over a batch of bursts.
 * NOT in the reference (core/protocol.py:300 "we do not do full channel decoding"): parity unpinned,
 * validated by encode -> decode round trips (tests/test_etsi_*.py).
 * d_soft [n][K3] int8 soft type-5 bits (+ = 1); channel 0 SCH/HD (124->216), 1 SCH/F (268->432),
 * 2 BSCH (60->120), 3 SCH/HU (92->168), all RCPC rate 2/3 with the 16-bit CRC; 4 TCH/4.8 (288->432, rate 292/432),
 * 5 TCH/2.4 (144->432, rate 148/432), both without a block code and with interleaving over N = 1 block;
 * scramb_init = ((colour | mnc<<6 | mcc<<20) << 2) | 3.
 * d_bits [n][40]: decoded type-1 bits packed MSB-first; d_crc_ok [n] (1 pass, 0 fail, 2 = the channel has no CRC);
 * d_metric [n] final path metric. */
int tb200_etsi_decode_batch(tb200_handle h, const int8_t *d_soft, int64_t n, int channel, uint32_t scramb_init,
                            uint8_t *d_bits, uint8_t *d_crc_ok, int32_t *d_metric, void *stream);

/* AACH (cl. 8.2.3.2): soft maximum-likelihood decoding of a (30, 14) block code over all 16 384 codewords.
 * d_soft [n][30] int8 soft type-5 bits; gen_rows: HOST array of the 14 generator rows, code bit j (first sent = 0) in word
 * bit 29 - j, row k belongs to information bit k + 1; d_bits [n] uint16: the 14 information bits, bit 1 = MSB (bit 13);
 * d_metric [n] the winning correlation, d_margin [n] (or NULL) its distance to the runner-up.  Ties go to the smaller
 * information word.  The standard's own generator matrix is NOT reproduced in this repository (no copy of EN 300 392-2
 * is available to it): tetraear_b200.etsi.rm3014_generator() builds a systematic shortened RM(2,5) of the same
 * parameters (30, 14, d = 8); pass the matrix of cl. 8.2.3.2 here for on-air signals. */
int tb200_etsi_aach_decode(tb200_handle h, const int8_t *d_soft, int64_t n, const uint32_t *gen_rows, uint32_t scramb_init,
                           uint16_t *d_bits, int32_t *d_metric, int32_t *d_margin, void *stream);

/* The same decoder for any block geometry: k2 = k1 + (crc ? 16 : 0) + 4 tail bits, k3 <= 432 type-3 bits, block
 * interleaver (k3, a), k2 <= 320; rate 0 = 2/3, 1 = 1/3, 2 = 292/432, 3 = 148/432 (EN 300 392-2 cl. 8.2.3.1.3-8.2.3.1.6:
 * type-3 bit j is mother bit 8 ((i-1) div t) + P(i - t ((i-1) div t)), i = j + (j-1) div {65, 35} for the last two). */
int tb200_etsi_decode_custom(tb200_handle h, const int8_t *d_soft, int64_t n, int k1, int k2, int k3, int a, int rate,
                             int crc, uint32_t scramb_init, uint8_t *d_bits, uint8_t *d_crc_ok, int32_t *d_metric,
                             void *stream);

/* `etsi` mode polyphase FIR channelizer (parity unpinned, SURVEY.md G1/G9: the reference has none):
 * wideband complex64 at fs = n_channels x 25 kHz -> n_channels rows at 50 kHz (2x oversampled);
 * row k is centred on k x 25 kHz (k >= n_channels/2 are the negative frequencies);
 * frame m = sum_n x[n] h[m D - n] exp(-j 2 pi k n / M), D = M/2, h = Kaiser-windowed sinc of 6 M taps.
 * d_out [n_channels][out_stride] complex64, n_frames <= out_stride frames written per row. */
int tb200_channelize(tb200_handle h, const void *d_iq_c64, int64_t n_samples, int n_channels, void *d_out_c64,
                     int64_t out_stride, int64_t n_frames, void *stream);

/* `etsi` mode physical layer (parity unpinned): per stream (one 25 kHz channel at
 * fs in [36, 72] kHz, i.e. 2..4 samples/symbol): RRC(0.35) matched filter + closed-loop timing
 * (Gardner + PI) + pi/4-DQPSK differential soft bits.  d_soft [n_streams][soft_stride] float, two per
 * symbol, + = 1, normalised by the running symbol power; d_n_soft = soft bits produced. */
int tb200_etsi_demod_batch(tb200_handle h, const void *d_iq_c64, int n_streams, int64_t n_samples, int64_t stride,
                           double fs, void *d_sym_c64, float *d_soft, int64_t soft_stride, int32_t *d_n_soft,
                           void *stream);
/* The same demodulator, time-parallel: every stream is cut into segments of >= 640 nominal symbols that run
 * concurrently (each re-locks its timing loop over 160 warm-up symbols before its first output) and are joined symbol
 * by symbol afterwards, so one long stream offers n_segments times the parallelism at 160/640 extra work.  Outputs as
 * tb200_etsi_demod_batch; they agree with it wherever the loop is locked (same symbols, instants within the loop's
 * jitter) and may differ in symbol count on a channel that carries only noise.  d_scratch: 16-byte aligned device
 * memory of tb200_etsi_demod_scratch_bytes(...) bytes; with less (or a stream shorter than two segments) the call
 * runs the sequential form. */
int64_t tb200_etsi_demod_scratch_bytes(int n_streams, int64_t n_samples, double fs, int want_symbols);
int tb200_etsi_demod_segmented(tb200_handle h, const void *d_iq_c64, int n_streams, int64_t n_samples, int64_t stride,
                               double fs, void *d_sym_c64, float *d_soft, int64_t soft_stride, int32_t *d_n_soft,
                               void *d_scratch, int64_t scratch_bytes, void *stream);
/* normal-burst training-sequence search on the soft bits: d_metric [n_streams][soft_stride/2] scratch
 * (normalised correlation per symbol position, sign = which of the two normal sequences); for every
 * burst found the two 216-bit blocks BKN1 | BKN2 are quantised to int8 into d_blocks
 * [n_streams][max_bursts][432] (input of tb200_etsi_decode_batch, channel 0) and its first bit
 * position into d_burst_pos. */
int tb200_etsi_find_bursts(tb200_handle h, const float *d_soft, int n_streams, int64_t soft_stride,
                           const int32_t *d_n_soft, double threshold, float *d_metric, int8_t *d_blocks,
                           int32_t *d_burst_pos, int32_t *d_n_bursts, int max_bursts, void *stream);
/* Two bit ranges [off_a, off_a + len_a) + [off_b, off_b + len_b) of every burst found by tb200_etsi_find_bursts /
 * tb200_etsi_find_sync_bursts, as int8 soft values (the blocks' scale): d_out [n_streams][max_bursts][len_a + len_b],
 * zero beyond a stream's burst count.  (230, 14) + (266, 16) is the broadcast block of a normal downlink burst = the 30
 * AACH bits, input of tb200_etsi_aach_decode. */
int tb200_etsi_extract_bits(tb200_handle h, const float *d_soft, int n_streams, int64_t soft_stride,
                            const int32_t *d_n_soft, const int32_t *d_burst_pos, const int32_t *d_n_bursts, int max_bursts,
                            int off_a, int len_a, int off_b, int len_b, int8_t *d_out, void *stream);

/* Search for any training sequence of EN 300 392-2 cl. 9.4.4.3 in the demodulator's soft bits: sequence 0 = normal n,
 * 1 = normal p, 2 = tail q (22 bits; in a continuous downlink q(1..10) closes a burst and q(11..22) opens the next: a hit
 * at bit p puts the slot boundary at p + 10), 3 = extended x (30 bits, uplink control bursts), 4 = synchronisation y (38).
 * Metric = clipped soft correlation / sum |soft| at every symbol position (d_metric [n_streams][soft_stride / 2],
 * scratch + output); a hit is the first position >= threshold refined to the maximum of the next 4 positions, after which
 * min_gap_symbols positions are skipped.  d_pos / d_corr [n_streams][max_hits]: bit offset and metric of each hit. */
int tb200_etsi_find_sequence(tb200_handle h, const float *d_soft, int n_streams, int64_t soft_stride,
                             const int32_t *d_n_soft, int sequence, double threshold, int min_gap_symbols,
                             float *d_metric, int32_t *d_pos, float *d_corr, int32_t *d_n_hits, int max_hits, void *stream);

/* feed-forward carrier-offset correction of the soft bits in place: theta = (arg sum d^4 - pi) / 4 over each stream's
 * differential symbols d (unambiguous for |df| < 2250 Hz), every soft pair rotated by -theta; d_theta [n_streams]
 * (radians per symbol; df = theta * 18000 / 2 pi) or NULL.  Run between tb200_etsi_demod_batch and the searches. */
int tb200_etsi_afc(tb200_handle h, float *d_soft, int n_streams, int64_t soft_stride, const int32_t *d_n_soft,
                   float *d_theta, void *stream);
/* synchronisation-burst search: normalised soft correlation with the 38-bit sync training sequence y at every symbol
 * position (d_metric scratch as above); for every sync burst found the 120 SB1 soft bits go to d_sb1
 * [n_streams][max_bursts][120] (input of tb200_etsi_decode_batch, channel 2 = BSCH, scrambling init 3: its 60 type-1
 * bits carry colour code, MCC and MNC, i.e. the scrambling init of every other channel) and, if d_bkn2 != NULL, the 216
 * BKN2 soft bits to d_bkn2 [n_streams][max_bursts][216]. */
int tb200_etsi_find_sync_bursts(tb200_handle h, const float *d_soft, int n_streams, int64_t soft_stride,
                                const int32_t *d_n_soft, double threshold, float *d_metric, int8_t *d_sb1, int8_t *d_bkn2,
                                int32_t *d_burst_pos, int32_t *d_n_bursts, int max_bursts, void *stream);

/* ---- waterfall PSD + AFC statistics (tetraear/ui/modern.py:1926-2029, :923-947) ---------
 * frame i = Hann(n_fft) * iq[i*frame_stride : +n_fft]; power = 20 log10(|fftshift FFT|/n_fft + 1e-20)
 * d_power    [n_frames][n_fft] float32 dB (or NULL)
 * d_smoothed [n_frames][n_fft] float32: EMA in the dB domain, S = S(1-a) + P a (or NULL)
 * d_ema_state [n_fft] double, carried across calls; *first* != 0 -> first frame initialises it
 * d_stats    [n_frames][8] double: signal_power, peak_power, peak_idx, peak_freq_offset,
 *            noise_floor, snr, signal_present, afc_offset   (modern.py:1956-2028) */
int tb200_psd_batch(tb200_handle h, const void *d_iq_c64, int64_t n_frames, int64_t frame_stride,
                    int n_fft, double sample_rate, double ema_alpha, int first, float *d_power,
                    float *d_smoothed, double *d_ema_state, double *d_stats, void *stream);

/* ---- per-stage device timing (bench.py's roofline): when enabled every entry point brackets
 * its stages with CUDA events on the launching stream; tb200_profile_read synchronises them,
 * sums the elapsed milliseconds per stage and clears the list.
 * stages: 0 decimator, 1 channel filter (NCO + Butterworth), 2 timing pick + slicer,
 *         3 bits + sync correlation, 4 sync select + frame/burst/CRC, 5 burst batch, 6 PSD, 7 etsi */
int tb200_profile_enable(tb200_handle h, int on);
int tb200_profile_read(tb200_handle h, double ms[8], int64_t calls[8]);

/* ---- reference-mode channel bank (BASELINE configs[3]) ----------------------------------
 * d_out[k][j], k = 0 .. n_channels-1, j = 0 .. ceil(n/q)-1  ==  scipy.signal.decimate(frequency_shift(x, f_k, fs), q)[j]
 * with f_k = k fs / n_channels for k < n_channels/2, (k - n_channels) fs / n_channels above (DFT bin order), i.e. what
 * SignalProcessor.frequency_shift + the decimation step of SignalProcessor.process compute for EVERY channel of a
 * wideband chunk (tetraear/signal/processor.py:85-100, :245-257; complex128 in the reference, so the float64 design),
 * with the input read once: polyphase fold of the zero-phase response + size-n_channels DFT, then scipy's chunk-end
 * terms (odd extension, steady-state continuation, backward-pass transient) per channel in float64.  complex64 out,
 * ~1e-6 of full scale from the reference.  Feed d_out to tb200_process_batch of a handle created for the decimated
 * rate fs/q (no further decimation there) to get dibits / symbols per channel.  The handle needs max_channels >= n_channels. */
int tb200_refchan_batch(tb200_handle h, const void *d_iq_c64, int64_t n_samples, int n_channels, int q, double fs,
                        void *d_out_c64, int64_t out_stride, void *stream);

/* ---- debugging aid: synchronous copy of an internal workspace region to the host ---------
 * region 0 = decimated chunks, 1 = channel-filtered complex128, 2 / 3 = forward-pass scratches,
 * 4 = the etsi demodulator's tap table [32 fractional phases][32 taps] float32 (valid after the first demodulation) */
int tb200_debug_read(tb200_handle h, int region, void *h_dst, size_t nbytes);
/* host only (no device needed): the weight matrix of the tensor-core decimator for factor q as K x 48 doubles
 * (row-major), its TF32 split in the shared-memory layout ([K/4][48][4] floats each) and the carry constants (flat
 * floats: lampow[5][4][2], lanepow[32][4][2], cF[3][4][2], cB[3][4][2], initF[4][2], initB[4][2]); tests emulate the kernel
 * with them on the CPU (tests/test_tc_design_cpu.py).  Any output pointer may be NULL. */
/* profiling aid: per-phase cycle counters of the tensor-core decimator (thread 0 of CTA 0); enable != 0 arms them for the
 * following launches on this handle; out (may be NULL) receives and clears the 16 counters (see dsp_tc.cu) */
int tb200_debug_tc_timing(tb200_handle h, int enable, int64_t out[16]);
int tb200_debug_tc_design(int q, double *B, int b_capacity, float *bhi, float *blo, int split_capacity, float *params,
                          int params_capacity, int *K_out);

/* ---- scanner heuristics, batched over dwells ----------------------------------------------
 * Replaces the per-sample Python loops of TetraSignalDetector.calculate_power / detect_tetra_modulation /
 * detect_sync_pattern / check_power_stability (tetraear/signal/scanner.py:42-55, :57-98, :100-147, :204-233).
 * The device returns the sums and integer counts of each dwell; the caller finishes the scalar formulas
 * (10 log10(sum/n + 1e-10), matches/n_steps > 0.4, sync_best/31 > 0.75, std(window dB) < 10). */
typedef struct tb200_detect_stats {
    double power_sum;        /* sum |x|^2 over the dwell */
    double window_sum[8];    /* sum |x|^2 over window w = samples [w*window_size, (w+1)*window_size) */
    double max_abs;          /* max |x| */
    int64_t n;               /* samples per dwell */
    int32_t window_size;     /* n / num_windows */
    int32_t n_steps;         /* phase steps examined by detect_tetra_modulation (0 when n < 1000) */
    int32_t matches;         /* ... of which within pi/8 of a multiple of pi/4 in [-pi, 3pi/4] */
    int32_t n_sync_bits;     /* bits of detect_sync_pattern's stream (0 when fewer than 100 "symbols") */
    int32_t sync_best;       /* best match count against the 31-bit pattern over range(n_sync_bits - 31) */
    int32_t reserved;
} tb200_detect_stats;        /* 112 bytes */

/* d_iq: [n_dwells][stride] complex (dtype C64 or C128, arithmetic follows the dtype like numpy's), n samples used
 * per dwell, fs = the detector's sample_rate (sets the sync stride int(fs/18000/10)), num_windows 0..8.
 * d_out: device array of n_dwells records.  Needs a handle with max_chunks*max_chunk_len >= n_dwells*n. */
int tb200_detect_batch(tb200_handle h, const void *d_iq, int dtype, int n_dwells, int64_t n, int64_t stride, double fs,
                       int num_windows, tb200_detect_stats *d_out, void *stream);

/* ---- codec-frame builder -------------------------------------------------------------------
 * 690 int16 per frame: [0] = 0x6B21, 432 soft values in 1..114, 116..229, 231..344, 346..435.
 *  TB200_VOICE_SYMBOLS  ui/modern.py:2309-2423 (_extract_voice_slot_from_symbols): d_src = dibit streams
 *                       [n_streams][src_stride] with d_n_src[stream] valid dibits, frame f lives in stream
 *                       d_frame_stream[f] (NULL = stream 0) at bit position d_frame_pos[f]; slot = dibits[pos/3 .. +255)
 *  TB200_VOICE_BITS     continuous_capture.py:100-124: d_src = bit rows [n_frames][src_stride], d_n_src[f] = row length
 *                       (NULL = src_stride); the first 432 bits -> +-127
 *  TB200_VOICE_PARAMS   ui/modern.py:2143-2200 (fallback parameter packing), same source as BITS
 *  TB200_VOICE_SOFT     d_src = float rows [n_frames][src_stride] of soft bits (+ = 1): clamp(rint(scale * s), +-127)
 * d_valid[f] = 0 (and an all-zero block) where the reference produces no frame. */
#define TB200_VOICE_SYMBOLS 0
#define TB200_VOICE_BITS 1
#define TB200_VOICE_PARAMS 2
#define TB200_VOICE_SOFT 3
int tb200_voice_blocks(tb200_handle h, int mode, const void *d_src, int64_t src_stride, const int32_t *d_n_src,
                       const int32_t *d_frame_stream, const int32_t *d_frame_pos, int64_t n_frames, double scale,
                       int16_t *d_blocks, uint8_t *d_valid, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* TETRA_B200_H */
