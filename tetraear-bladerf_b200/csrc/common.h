// common.h — shared host-side declarations of libtetra_b200 (internal).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>
#include "../../include/tetra_b200.h"

namespace tb200 {

int fail(int code, const char *fmt, ...);

#define TB_CUDA(expr)                                                                          \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess)                                                                 \
            return tb200::fail(TB200_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                               __FILE__, __LINE__);                                            \
    } while (0)

#define TB_LAUNCH_CHECK(h)                                                                     \
    do {                                                                                       \
        (h)->launches++;                                                                       \
        cudaError_t _e = cudaGetLastError();                                                   \
        if (_e != cudaSuccess)                                                                 \
            return tb200::fail(TB200_E_CUDA, "kernel launch failed: %s (%s:%d)",               \
                               cudaGetErrorString(_e), __FILE__, __LINE__);                    \
    } while (0)

// Filter constants for one (sample_rate) setting, as kernels take them (by value).
struct DecimDesign {
    float sos32[24];
    float zi32[8];
    double sos64[24];
    double zi64[8];
};
struct ChanDesign {
    double b[5];
    double a[5];
    double zi[4];
};

}  // namespace tb200

#define TB200_N_STAGES 8
enum { TB_STAGE_DECIM = 0, TB_STAGE_CHAN = 1, TB_STAGE_SYMBOLS = 2, TB_STAGE_BITS_SYNC = 3, TB_STAGE_DECODE = 4,
       TB_STAGE_BURST = 5, TB_STAGE_PSD = 6, TB_STAGE_ETSI = 7 };

struct tb200_handle_s {
    tb200_config cfg;
    int64_t launches;
    // optional per-stage CUDA-event timing (tb200_profile_*): pairs recorded on the launching stream
    int psd_f32;                 // PSD: float32 FFT + dB (dense config mode) instead of float64
    int etsi_seg_lanes;          // tb200_etsi_demod_segmented: 0 = two-stage form (default), 8 = eight-lane kernel
    int etsi_open_loop;          // etsi demodulator: timing loop gains 0 (fixed symbol period; validation of the matched filter)
    int psd_avg;                 // PSD: linear (Welch) average of this many consecutive frames per output frame (>= 1)
    int chan_f32;                // FAST mode: run the channel filter's recurrences in float32 (tb200_set_option)
    int decim_tc;                // FAST mode: tensor-core decimator (dsp_tc.cu) where its geometry rules hold (default on)
    int profiling;
    std::vector<cudaEvent_t> *prof_ev;      // flat: start, stop, start, stop ...
    std::vector<int> *prof_stage;
    tb200_plan plan;             // for (cfg.sample_rate, cfg.max_chunk_len)
    // cached designs
    int dec_q;                   // q the decimator design below is for (0 = none)
    tb200::DecimDesign dec;
    double chan_rate, chan_bw;   // (rate, bandwidth) the channel design below is for
    tb200::ChanDesign chan;
    // workspace: one device allocation, carved
    void *ws;
    size_t ws_bytes;
    void *fwd;        size_t fwd_bytes;       // decimator forward scratch [time][lane]
    void *dec_out;    size_t dec_bytes;       // decimated chunks (complex f32 / f64)
    double *chan_fwd; size_t chan_bytes;      // channel-filter forward scratch [time][lane]
    double *filtered; size_t filt_bytes;      // complex128 [chunk][n_dec]
    int32_t *n_sym;                           // [chunks] symbols per chunk
    double *phase_pw; size_t phase_bytes;     // per-(chunk, tile) phase power partial sums (FAST mode)
    // decode
    int max_streams; long bits_stride; int max_frames;
    uint8_t *bits, *mapped, *m1, *m2;
    int32_t *sync_pos, *sync_cnt;             // [streams][max_frames], [2]
    // psd
    double *psd_tables; int psd_tables_n;
    float *chan_proto, *chan_tw; int chan_M;   // channelizer prototype [6 M], twiddles, M they are for
    double etsi_taps_fs;                      // fs the tap table below was built for (0 = none)
    float *etsi_taps;                         // [32 phases][32 taps] interpolating matched filter
    float *etsi_itab;                         // [32 phases][8 taps] fractional-delay interpolator of the two-stage demodulator
    float *tc_bmat; size_t tc_bmat_bytes;     // tensor-core decimator weights (Bhi, Blo in the canonical layout)
    int tc_key_q; float tc_key_sos[24];       // design the weights above were built for (q = 0: none)
    long long *tc_timing;                     // phase counters of the tensor-core decimator (profiling aid) or null
    void *rc_tab; size_t rc_tab_bytes;        // reference-mode channel bank: folded prototype, twiddles, g, channel frequencies
    int rc_M, rc_q; long rc_n; double rc_fs;  // what the tables above were built for (rc_M = 0: none)
    // staging for the *_host entry points
    void *stage_in;   size_t stage_in_bytes;
    void *stage_sc16;                         // int16 IQ staging of the host entry point
    uint8_t *stage_dibits; void *stage_symbols; int32_t *stage_counts; double *stage_freq;
    size_t stage_out_stride;
    uint8_t *stage_burst_in, *stage_burst_rec, *stage_burst_full; int64_t *stage_stats;
    // one-shot result region of tb200_process_decode_host / tb200_decode_batch_host: [header: n_dibits[C], phase[C],
    // n_frames[C], stats[3]] [dibits C x stride] [frames C x max_frames x 156] [symbols C x stride x 16], copied to the
    // pinned mirror `pin` with ONE device-to-host copy (0 bytes: batch too large, piecewise copies instead)
    char *oshot; size_t oshot_bytes; size_t oshot_streams, oshot_stride, oshot_frames;
    uint8_t *stage_dec_in; int32_t *stage_dec_cnt;
    char *pin; size_t pin_bytes;
};
