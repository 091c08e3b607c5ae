// api.cu — the C-ABI of libtetra_b200 (include/tetra_b200.h): handle, workspace, entry points.
#include <cstdarg>
#include <cstring>
#include <new>
#include "common.h"
#include "kernels.h"

namespace tb200 {

static thread_local char g_err[512] = "";

int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

// Entry points run on the handle's device whatever the caller's current device is (several handles on different
// GPUs may live in one process); the caller's device is restored on exit.
struct DevGuard {
    int prev = -1;
    bool switched = false;
    explicit DevGuard(tb200_handle h)
    {
        if (!h) return;
        if (cudaGetDevice(&prev) == cudaSuccess && prev != h->cfg.device) switched = cudaSetDevice(h->cfg.device) == cudaSuccess;
    }
    ~DevGuard() { if (switched) cudaSetDevice(prev); }
};

int ensure_crc_tables(int device);
int launch_decimate_fast(tb200_handle h, const void *x, int n_chunks, long n, long chunk_stride, int q,
                         const DecimDesign &d, void *dec_out, long n_dec, cudaStream_t st);
int fast_decimator_supported(int q);
bool tc_decimator_usable(const void *x, long n, long chunk_stride, int q);
int launch_decimate_tc(tb200_handle h, const void *x, int n_chunks, long n, long chunk_stride, int q, const DecimDesign &d,
                       void *dec_out, long n_dec, cudaStream_t st);
int launch_channel_fast(tb200_handle h, const void *x, int dtype, int n_chunks, long n, long in_stride,
                        const double *freq_off, double fs, double bandwidth, double *filtered, long out_stride,
                        double *phase_partial, double *phase_tail, int sps, int step, cudaStream_t st);
int chan_fast_tiles(tb200_handle h, int dtype, long n);

static size_t align_up(size_t v) { return (v + 255) & ~(size_t)255; }

struct Carver {
    char *base;
    size_t off;
    template <typename T> T *take(size_t bytes)
    {
        T *p = base ? reinterpret_cast<T *>(base + off) : nullptr;
        off += align_up(bytes);
        return p;
    }
};

// lay out (or, with base == nullptr, just size) the workspace of a handle
static size_t carve(tb200_handle h, char *base)
{
    Carver c{base, 0};
    const tb200_config &cfg = h->cfg;
    const size_t C = (size_t)cfg.max_chunks, L = (size_t)cfg.max_chunk_len;
    const tb200_plan &P = h->plan;
    const size_t ND = (size_t)P.n_dec;
    if (C && L) {
        h->fwd_bytes = C * 2 * (L + 54) * sizeof(float);
        h->fwd = c.take<char>(h->fwd_bytes);
        h->dec_bytes = C * (L / 2 + 2) * 16;
        h->dec_out = c.take<char>(h->dec_bytes);
        h->chan_bytes = C * 2 * (ND + 30) * sizeof(double);
        h->chan_fwd = c.take<double>(h->chan_bytes);
        h->filt_bytes = C * ND * 16;
        h->filtered = c.take<double>(h->filt_bytes);
        h->n_sym = c.take<int32_t>(C * sizeof(int32_t));
        h->phase_bytes = C * (size_t)(ND / 1536 + 3) * TB200_MAX_PHASES * sizeof(double);
        h->phase_pw = c.take<double>(h->phase_bytes);
    } else {
        h->fwd = h->dec_out = nullptr; h->chan_fwd = h->filtered = nullptr; h->n_sym = nullptr;
        h->phase_pw = nullptr; h->phase_bytes = 0;
        h->fwd_bytes = h->dec_bytes = h->chan_bytes = h->filt_bytes = 0;
    }
    // decode: per stream a bit row of bits_stride bytes
    size_t S = (size_t)h->max_streams, B = (size_t)h->bits_stride;
    h->bits = c.take<uint8_t>(S * B);
    h->mapped = c.take<uint8_t>(S * B / 2);
    h->m1 = c.take<uint8_t>(S * B);
    h->m2 = c.take<uint8_t>(S * B);
    h->sync_pos = c.take<int32_t>(S * (size_t)h->max_frames * sizeof(int32_t));
    h->sync_cnt = c.take<int32_t>(64);
    size_t nf = (size_t)cfg.max_fft;
    h->psd_tables = c.take<double>((2 * nf + 2 * (nf / 2 + nf / 4 + 1)) * sizeof(double));
    h->etsi_taps = c.take<float>(32 * 32 * sizeof(float));
    h->etsi_itab = c.take<float>(32 * 8 * sizeof(float));
    h->tc_bmat_bytes = (C && L) ? 2 * 16 * 48 * 16 + 128 : 0;     // Bhi + Blo, up to 16 K-chunks x 48 columns x 16 bytes (+ 16 phase counters)
    h->tc_bmat = c.take<float>(h->tc_bmat_bytes);
    size_t MC = (size_t)cfg.max_channels;
    h->chan_proto = c.take<float>(MC * 6 * sizeof(float));
    h->chan_tw = c.take<float>((MC + 256) * 2 * sizeof(float));
    h->rc_tab_bytes = MC ? MC * 32 * sizeof(float) + (MC + 512) * 2 * sizeof(float) + (MC * 16 + 64) * sizeof(double) + MC * sizeof(double) + 16 * MC * 2 * sizeof(double) : 0;
    h->rc_tab = MC ? c.take<char>(h->rc_tab_bytes) : nullptr;
    // staging for *_host
    h->stage_in_bytes = C * L * 8;
    h->stage_in = c.take<char>(h->stage_in_bytes);
    h->stage_sc16 = c.take<char>(h->stage_in_bytes / 2);
    h->stage_out_stride = (size_t)P.max_symbols + 8;
    h->stage_dibits = c.take<uint8_t>(C * h->stage_out_stride);
    h->stage_symbols = c.take<char>(C * h->stage_out_stride * 16);
    h->stage_counts = c.take<int32_t>(2 * C * sizeof(int32_t) + 64);
    h->stage_freq = c.take<double>(C * sizeof(double) + 64);
    size_t NB = (size_t)cfg.max_bursts;
    h->stage_burst_in = c.take<uint8_t>(NB * 255);
    h->stage_burst_rec = c.take<uint8_t>(NB * 32);
    h->stage_burst_full = c.take<uint8_t>(NB * 64);
    h->stage_stats = c.take<int64_t>(64);
    // one-shot result region (small batches only: the class API works chunk by chunk)
    {
        const size_t SS = C ? C : S;
        const size_t stride = C ? h->stage_out_stride : B / 2;
        const size_t MF = 64 < (size_t)h->max_frames ? 64 : (size_t)h->max_frames;
        const size_t bytes = 256 + SS * 16 + SS * stride + 64 + SS * MF * sizeof(tb200_frame) + 64 + (C ? SS * stride * 16 : 0);
        if (bytes <= (1u << 20)) {
            h->oshot_bytes = bytes; h->oshot_streams = SS; h->oshot_stride = stride; h->oshot_frames = MF;
            h->oshot = c.take<char>(bytes);
        } else {
            h->oshot_bytes = 0; h->oshot = nullptr; h->oshot_streams = h->oshot_stride = h->oshot_frames = 0;
        }
        h->stage_dec_in = c.take<uint8_t>(S * (B / 2));
        h->stage_dec_cnt = c.take<int32_t>(S * sizeof(int32_t) + 64);
    }
    return c.off;
}

static int refresh_designs(tb200_handle h)
{
    int rc = tb200_plan_process(h->cfg.sample_rate, h->cfg.max_chunk_len, &h->plan);
    if (rc) return rc;
    h->dec_q = 0;
    h->chan_rate = h->chan_bw = -1.0;
    return TB200_OK;
}

static int get_decim(tb200_handle h, int q)
{
    if (h->dec_q == q) return TB200_OK;
    int rc = tb200_design_decimator(q, h->dec.sos64, h->dec.zi64, h->dec.sos32, h->dec.zi32);
    if (rc) return rc;
    h->dec_q = q;
    return TB200_OK;
}

static int get_chan(tb200_handle h, double bw, double rate)
{
    if (h->chan_rate == rate && h->chan_bw == bw) return TB200_OK;
    int rc = tb200_design_channel_filter(bw, rate, h->chan.b, h->chan.a, h->chan.zi);
    if (rc) return rc;
    h->chan_rate = rate; h->chan_bw = bw;
    return TB200_OK;
}

// RAII stage timer: records an event pair around a stage when profiling is on
struct StageTimer {
    tb200_handle h; cudaStream_t st; bool on;
    StageTimer(tb200_handle h_, int stage, cudaStream_t st_) : h(h_), st(st_), on(h_->profiling != 0)
    {
        if (!on) return;
        // events recorded into a graph capture cannot be read back with cudaEventElapsedTime: no timing while capturing
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(st, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone) { cudaGetLastError(); on = false; return; }
        cudaEvent_t a, b;
        if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) { on = false; return; }
        h->prof_ev->push_back(a); h->prof_ev->push_back(b); h->prof_stage->push_back(stage);
        cudaEventRecord(a, st);
    }
    ~StageTimer() { if (on) cudaEventRecord(h->prof_ev->back(), st); }
};

}  // namespace tb200

using namespace tb200;

extern "C" int tb200_profile_enable(tb200_handle h, int on)
{
    DevGuard dev_guard(h);
    if (!h) return fail(TB200_E_ARG, "tb200_profile_enable: null handle");
    if (!h->prof_ev) { h->prof_ev = new std::vector<cudaEvent_t>(); h->prof_stage = new std::vector<int>(); }
    h->profiling = on;
    return TB200_OK;
}

extern "C" int tb200_profile_read(tb200_handle h, double ms[TB200_N_STAGES], int64_t calls[TB200_N_STAGES])
{
    DevGuard dev_guard(h);
    if (!h || !ms || !calls) return fail(TB200_E_ARG, "tb200_profile_read: bad argument");
    for (int i = 0; i < TB200_N_STAGES; ++i) { ms[i] = 0.0; calls[i] = 0; }
    if (!h->prof_ev) return TB200_OK;
    for (size_t k = 0; k < h->prof_stage->size(); ++k) {
        cudaEvent_t a = (*h->prof_ev)[2 * k], b = (*h->prof_ev)[2 * k + 1];
        TB_CUDA(cudaEventSynchronize(b));
        float t = 0.f;
        TB_CUDA(cudaEventElapsedTime(&t, a, b));
        int sgi = (*h->prof_stage)[k];
        ms[sgi] += t; calls[sgi] += 1;
        cudaEventDestroy(a); cudaEventDestroy(b);
    }
    h->prof_ev->clear(); h->prof_stage->clear();
    return TB200_OK;
}

extern "C" const char *tb200_last_error(void) { return g_err; }
extern "C" int tb200_version(void) { return 100; }

extern "C" int tb200_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" int tb200_create(const tb200_config *cfg, tb200_handle *out)
{
    if (!cfg || !out) return fail(TB200_E_ARG, "tb200_create: null argument");
    if (cfg->max_chunks < 0 || cfg->max_chunk_len < 0 || !(cfg->sample_rate > 0))
        return fail(TB200_E_ARG, "tb200_create: bad capacity / sample rate");
    int ndev = tb200_device_count();
    if (ndev <= 0) return fail(TB200_E_NODEV, "tb200_create: no CUDA device visible (there is no CPU fallback)");
    if (cfg->device < 0 || cfg->device >= ndev) return fail(TB200_E_ARG, "tb200_create: device %d of %d", cfg->device, ndev);
    struct Restore { int prev = -1; Restore() { if (cudaGetDevice(&prev) != cudaSuccess) prev = -1; } ~Restore() { if (prev >= 0) cudaSetDevice(prev); } } restore;
    TB_CUDA(cudaSetDevice(cfg->device));
    tb200_handle h = new (std::nothrow) tb200_handle_s();
    if (!h) return fail(TB200_E_ALLOC, "tb200_create: out of host memory");
    memset(h, 0, sizeof(*h));
    h->cfg = *cfg;
    if (h->cfg.max_fft <= 0) h->cfg.max_fft = 4096;
    int rc = refresh_designs(h);
    if (rc) { delete h; return rc; }
    h->max_streams = cfg->max_chunks > 0 ? cfg->max_chunks : 1;
    long per_stream = h->plan.max_symbols + 8;
    if (per_stream < 4096) per_stream = 4096;
    if (cfg->max_stream_len > per_stream) per_stream = cfg->max_stream_len;
    h->bits_stride = 2 * ((per_stream + 127) / 128 * 128);
    h->max_frames = (int)(h->bits_stride / 250 + 2);
    h->psd_tables_n = 0;
    h->chan_M = 0;
    h->decim_tc = 0;
    h->rc_M = 0;
    h->tc_key_q = 0;
    h->ws_bytes = carve(h, nullptr);
    cudaError_t e = cudaMalloc(&h->ws, h->ws_bytes);
    if (e != cudaSuccess) {
        size_t want = h->ws_bytes;
        delete h;
        return fail(TB200_E_ALLOC, "tb200_create: cudaMalloc(%zu bytes) failed: %s", want, cudaGetErrorString(e));
    }
    carve(h, (char *)h->ws);
    rc = ensure_crc_tables(cfg->device);
    if (rc) { cudaFree(h->ws); delete h; return rc; }
    if (h->oshot_bytes) {                                   // pinned mirror of the one-shot result region
        if (cudaHostAlloc((void **)&h->pin, h->oshot_bytes, cudaHostAllocDefault) == cudaSuccess) h->pin_bytes = h->oshot_bytes;
        else { cudaGetLastError(); h->pin = nullptr; h->pin_bytes = 0; }
    }
    *out = h;
    return TB200_OK;
}

extern "C" int tb200_destroy(tb200_handle h)
{
    if (!h) return TB200_OK;
    DevGuard dev_guard(h);
    if (h->ws) cudaFree(h->ws);
    if (h->pin) cudaFreeHost(h->pin);
    if (h->prof_ev) {
        for (cudaEvent_t e : *h->prof_ev) cudaEventDestroy(e);
        delete h->prof_ev; delete h->prof_stage;
    }
    delete h;
    return TB200_OK;
}

extern "C" int tb200_set_sample_rate(tb200_handle h, double sample_rate)
{
    DevGuard dev_guard(h);
    if (!h || !(sample_rate > 0)) return fail(TB200_E_ARG, "tb200_set_sample_rate: bad argument");
    tb200_handle_s probe = *h;
    probe.cfg.sample_rate = sample_rate;
    int rc = refresh_designs(&probe);
    if (rc) return rc;
    probe.bits_stride = h->bits_stride; probe.max_frames = h->max_frames; probe.max_streams = h->max_streams;
    if (probe.plan.max_symbols + 8 > h->bits_stride / 2 || carve(&probe, nullptr) > h->ws_bytes)
        return fail(TB200_E_ARG, "tb200_set_sample_rate: %g Hz needs a larger workspace; create a new handle", sample_rate);
    probe.ws = h->ws; probe.ws_bytes = h->ws_bytes; probe.launches = h->launches;
    // every table carved behind the plan-dependent regions moves with n_dec: forget what was uploaded into them
    probe.psd_tables_n = 0; probe.etsi_taps_fs = 0.0; probe.chan_M = 0; probe.tc_key_q = 0; probe.tc_timing = nullptr; probe.rc_M = 0;
    carve(&probe, (char *)h->ws);
    *h = probe;
    return TB200_OK;
}

extern "C" int tb200_set_option(tb200_handle h, int option, int value)
{
    DevGuard dev_guard(h);
    if (!h) return fail(TB200_E_ARG, "tb200_set_option: null handle");
    switch (option) {
    case TB200_OPT_CHAN_F32: h->chan_f32 = value ? 1 : 0; return TB200_OK;
    case TB200_OPT_PSD_F32: h->psd_f32 = value ? 1 : 0; return TB200_OK;
    case TB200_OPT_DECIM_TC: h->decim_tc = value < 0 ? 0 : (value > 2 ? 2 : value); return TB200_OK;
    case TB200_OPT_PSD_AVG: h->psd_avg = value > 1 ? value : 1; return TB200_OK;
    case TB200_OPT_ETSI_OPEN_LOOP: h->etsi_open_loop = value ? 1 : 0; return TB200_OK;
    case TB200_OPT_ETSI_SEG_LANES: h->etsi_seg_lanes = value == 8 ? 8 : 0; return TB200_OK;
    default: return fail(TB200_E_ARG, "tb200_set_option: unknown option %d", option);
    }
}

extern "C" size_t tb200_workspace_bytes(tb200_handle h) { return h ? h->ws_bytes : 0; }
extern "C" int64_t tb200_launch_count(tb200_handle h) { return h ? h->launches : 0; }

// ------------------------------------------------------------------------- process
// process() of a batch; with d_frames the decode() of every chunk follows: fused into the last launch where the chunk
// fits the one-launch decoder (tb200_process_decode_batch), otherwise as the separate decode path
static int process_impl(tb200_handle h, const void *d_iq, int dtype, int n_chunks, int chunk_len,
                        int64_t chunk_stride, const double *d_freq_off, uint8_t *d_dibits,
                        void *d_symbols, int out_stride, int32_t *d_n_dibits, int32_t *d_phase,
                        tb200_frame *d_frames, int max_frames, int32_t *d_n_frames, int64_t *d_stats, void *stream)
{
    DevGuard dev_guard(h);
    if (!h || !d_n_dibits || n_chunks < 0 || chunk_len < 0) return fail(TB200_E_ARG, "tb200_process_batch: bad argument");
    if (dtype != TB200_DTYPE_C64 && dtype != TB200_DTYPE_C128 && dtype != TB200_DTYPE_SC16)
        return fail(TB200_E_ARG, "tb200_process_batch: dtype");
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == TB200_DTYPE_SC16 && n_chunks > 0 && chunk_len > 0) {
        // int16 IQ: convert into the complex64 staging area, then this is the ordinary complex64 path
        if (!d_iq) return fail(TB200_E_ARG, "tb200_process_batch: null buffer");
        if (n_chunks > h->cfg.max_chunks || chunk_len > h->cfg.max_chunk_len || chunk_stride != chunk_len)
            return fail(TB200_E_ARG, "tb200_process_batch: SC16 input must be dense and within the handle capacity");
        int rc = launch_sc16_to_c64(h, d_iq, (long)n_chunks * chunk_len, h->stage_in, st);
        if (rc) return rc;
        d_iq = h->stage_in;
        dtype = TB200_DTYPE_C64;
    }
    if (n_chunks == 0) return TB200_OK;
    if (n_chunks > h->cfg.max_chunks || chunk_len > h->cfg.max_chunk_len)
        return fail(TB200_E_ARG, "tb200_process_batch: %d x %d exceeds the handle capacity %d x %d", n_chunks, chunk_len,
                    h->cfg.max_chunks, h->cfg.max_chunk_len);
    if (chunk_len == 0) {                                            // processor.py:239-241
        TB_CUDA(cudaMemsetAsync(d_n_dibits, 0, sizeof(int32_t) * n_chunks, st));
        if (d_phase) TB_CUDA(cudaMemsetAsync(d_phase, 0, sizeof(int32_t) * n_chunks, st));
        return TB200_OK;
    }
    if (!d_iq || !d_dibits) return fail(TB200_E_ARG, "tb200_process_batch: null buffer");
    if (chunk_stride < chunk_len) return fail(TB200_E_ARG, "tb200_process_batch: chunk_stride < chunk_len");
    tb200_plan P;
    int rc = tb200_plan_process(h->cfg.sample_rate, chunk_len, &P);
    if (rc) return rc;
    if (out_stride < P.max_symbols) return fail(TB200_E_ARG, "tb200_process_batch: out_stride %d < %d symbols", out_stride, P.max_symbols);
    const size_t esz = dtype == TB200_DTYPE_C64 ? 4 : 8;
    const void *x = d_iq;
    long in_stride = (long)chunk_stride;
    int cur_dtype = dtype;
    if (P.q > 1) {
        rc = get_decim(h, P.q);
        if (rc) return rc;
        if ((size_t)n_chunks * P.n_dec * 2 * esz > h->dec_bytes)
            return fail(TB200_E_ARG, "tb200_process_batch: decimated batch exceeds workspace");
        StageTimer tm(h, TB_STAGE_DECIM, st);
        if (h->cfg.mode == TB200_MODE_FAST && dtype == TB200_DTYPE_C64 && h->decim_tc &&
            tc_decimator_usable(d_iq, chunk_len, (long)chunk_stride, P.q)) {
            rc = launch_decimate_tc(h, d_iq, n_chunks, chunk_len, (long)chunk_stride, P.q, h->dec, h->dec_out, P.n_dec, st);
        } else if (h->cfg.mode == TB200_MODE_FAST && dtype == TB200_DTYPE_C64 && fast_decimator_supported(P.q)) {
            rc = launch_decimate_fast(h, d_iq, n_chunks, chunk_len, (long)chunk_stride, P.q, h->dec, h->dec_out, P.n_dec, st);
        } else {
            if ((size_t)n_chunks * 2 * (chunk_len + 54) * esz > h->fwd_bytes)
                return fail(TB200_E_ARG, "tb200_process_batch: batch exceeds the decimator scratch (complex128 input needs 2x)");
            rc = launch_decimate_exact(h, d_iq, dtype, n_chunks, chunk_len, (long)chunk_stride, P.q, h->dec, h->fwd,
                                       h->dec_out, P.n_dec, st);
        }
        if (rc) return rc;
        x = h->dec_out;
        in_stride = P.n_dec;
    }
    if ((size_t)n_chunks * 2 * (P.n_dec + 30) * sizeof(double) > h->chan_bytes || (size_t)n_chunks * P.n_dec * 16 > h->filt_bytes)
        return fail(TB200_E_ARG, "tb200_process_batch: batch exceeds the channel-filter workspace");
    rc = get_chan(h, 25000.0, P.rate);
    if (rc) return rc;
    double *ph_partial = nullptr, *ph_tail = nullptr;
    int ph_tiles = 0;
    {
        StageTimer tm(h, TB_STAGE_CHAN, st);
        if (h->cfg.mode == TB200_MODE_FAST && P.filtered && chan_fast_supported(25000.0, P.rate)) {
            int tiles = chan_fast_tiles(h, (P.q > 1) ? cur_dtype : dtype, P.n_dec);
            if (P.sps > 1 && P.sps <= TB200_MAX_PHASES &&
                (size_t)n_chunks * (tiles + 1) * TB200_MAX_PHASES * sizeof(double) <= h->phase_bytes) {
                ph_partial = h->phase_pw;
                ph_tail = nullptr;          // the channel filter applies its right-edge term itself: no separate tail sums
                ph_tiles = tiles;
            }
            rc = launch_channel_fast(h, x, (P.q > 1) ? cur_dtype : dtype, n_chunks, P.n_dec, in_stride, d_freq_off, P.rate,
                                     25000.0, h->filtered, P.n_dec, ph_partial, ph_tail, P.sps, P.phase_step, st);
        }
        else
            rc = launch_channel_exact(h, x, cur_dtype, n_chunks, P.n_dec, in_stride, d_freq_off, P.rate, h->chan, P.filtered,
                                      h->chan_fwd, h->filtered, P.n_dec, st);
    }
    if (rc) return rc;
    if (d_frames && out_stride <= 6000 && max_frames <= 64) {
        StageTimer tm(h, TB_STAGE_SYMBOLS, st);            // timing pick + slicer + decode in one launch
        return launch_tail_fused(h, h->filtered, n_chunks, P.n_dec, P.n_dec, P.sps, P.phase_step, (double *)d_symbols, d_dibits,
                                 out_stride, d_n_dibits, d_phase, ph_partial, ph_tail, ph_tiles, d_frames, max_frames, d_n_frames,
                                 d_stats, st);
    }
    {
        StageTimer tm(h, TB_STAGE_SYMBOLS, st);
        rc = launch_symbols(h, h->filtered, n_chunks, P.n_dec, P.n_dec, P.sps, P.phase_step, (double *)d_symbols, d_dibits,
                            out_stride, d_n_dibits, nullptr, d_phase, st, ph_partial, ph_tail, ph_tiles);
    }
    if (rc || !d_frames) return rc;
    return tb200_decode_batch(h, d_dibits, n_chunks, out_stride, d_n_dibits, d_frames, max_frames, d_n_frames, d_stats, stream);
}

extern "C" int tb200_process_batch(tb200_handle h, const void *d_iq, int dtype, int n_chunks, int chunk_len,
                                   int64_t chunk_stride, const double *d_freq_off, uint8_t *d_dibits,
                                   void *d_symbols, int out_stride, int32_t *d_n_dibits, int32_t *d_phase,
                                   void *stream)
{
    return process_impl(h, d_iq, dtype, n_chunks, chunk_len, chunk_stride, d_freq_off, d_dibits, d_symbols, out_stride, d_n_dibits,
                        d_phase, nullptr, 0, nullptr, nullptr, stream);
}

extern "C" int tb200_process_decode_batch(tb200_handle h, const void *d_iq, int dtype, int n_chunks, int chunk_len,
                                          int64_t chunk_stride, const double *d_freq_off, uint8_t *d_dibits,
                                          void *d_symbols, int out_stride, int32_t *d_n_dibits, int32_t *d_phase,
                                          tb200_frame *d_frames, int max_frames, int32_t *d_n_frames, int64_t *d_stats,
                                          void *stream)
{
    if (!d_frames || !d_n_frames || max_frames <= 0) return fail(TB200_E_ARG, "tb200_process_decode_batch: bad argument");
    if (n_chunks > 0 && chunk_len == 0) {
        cudaError_t e = cudaMemsetAsync(d_n_frames, 0, sizeof(int32_t) * n_chunks, (cudaStream_t)stream);
        if (e != cudaSuccess) return fail(TB200_E_CUDA, "tb200_process_decode_batch: %s", cudaGetErrorString(e));
    }
    return process_impl(h, d_iq, dtype, n_chunks, chunk_len, chunk_stride, d_freq_off, d_dibits, d_symbols, out_stride, d_n_dibits,
                        d_phase, d_frames, max_frames, d_n_frames, d_stats, stream);
}

extern "C" int tb200_process_batch_host(tb200_handle h, const void *h_iq, int dtype, int n_chunks, int chunk_len,
                                        const double *h_freq_off, uint8_t *h_dibits, void *h_symbols, int out_stride,
                                        int32_t *h_n_dibits, int32_t *h_phase, void *stream)
{
    DevGuard dev_guard(h);
    if (!h || !h_n_dibits) return fail(TB200_E_ARG, "tb200_process_batch_host: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    if (n_chunks <= 0) return TB200_OK;
    if (n_chunks > h->cfg.max_chunks || chunk_len > h->cfg.max_chunk_len)
        return fail(TB200_E_ARG, "tb200_process_batch_host: exceeds the handle capacity");
    const size_t esz = dtype == TB200_DTYPE_C64 ? 8 : (dtype == TB200_DTYPE_SC16 ? 4 : 16);
    size_t in_bytes = (size_t)n_chunks * chunk_len * esz;
    void *stage = dtype == TB200_DTYPE_SC16 ? h->stage_sc16 : h->stage_in;
    if (in_bytes > (dtype == TB200_DTYPE_SC16 ? h->stage_in_bytes / 2 : h->stage_in_bytes))
        return fail(TB200_E_ARG, "tb200_process_batch_host: input exceeds staging (complex128 needs 2x)");
    if ((size_t)out_stride > h->stage_out_stride) return fail(TB200_E_ARG, "tb200_process_batch_host: out_stride too large");
    if (in_bytes) TB_CUDA(cudaMemcpyAsync(stage, h_iq, in_bytes, cudaMemcpyHostToDevice, st));
    double *d_f = nullptr;
    if (h_freq_off) {
        TB_CUDA(cudaMemcpyAsync(h->stage_freq, h_freq_off, sizeof(double) * n_chunks, cudaMemcpyHostToDevice, st));
        d_f = h->stage_freq;
    }
    int32_t *d_cnt = h->stage_counts, *d_ph = h->stage_counts + n_chunks;
    int rc = tb200_process_batch(h, stage, dtype, n_chunks, chunk_len, chunk_len, d_f, h->stage_dibits,
                                 h_symbols ? h->stage_symbols : nullptr, out_stride, d_cnt, d_ph, stream);
    if (rc) return rc;
    if (h_dibits) TB_CUDA(cudaMemcpyAsync(h_dibits, h->stage_dibits, (size_t)n_chunks * out_stride, cudaMemcpyDeviceToHost, st));
    if (h_symbols) TB_CUDA(cudaMemcpyAsync(h_symbols, h->stage_symbols, (size_t)n_chunks * out_stride * 16, cudaMemcpyDeviceToHost, st));
    TB_CUDA(cudaMemcpyAsync(h_n_dibits, d_cnt, sizeof(int32_t) * n_chunks, cudaMemcpyDeviceToHost, st));
    if (h_phase) TB_CUDA(cudaMemcpyAsync(h_phase, d_ph, sizeof(int32_t) * n_chunks, cudaMemcpyDeviceToHost, st));
    TB_CUDA(cudaStreamSynchronize(st));
    return TB200_OK;
}

// Layout of the one-shot result region for (n streams, stride, max_frames): offsets of its parts
struct OneShot { size_t cnt, phase, nfr, stats, dibits, frames, symbols, end_no_sym, end; };
static OneShot oneshot_layout(size_t n, size_t stride, size_t mf, bool with_symbols)
{
    OneShot o;
    o.cnt = 0; o.phase = n * 4; o.nfr = 2 * n * 4; o.stats = (3 * n * 4 + 15) & ~(size_t)15;
    o.dibits = o.stats + 32;
    o.frames = (o.dibits + n * stride + 15) & ~(size_t)15;
    o.end_no_sym = o.frames + n * mf * sizeof(tb200_frame);
    o.symbols = (o.end_no_sym + 15) & ~(size_t)15;
    o.end = with_symbols ? o.symbols + n * stride * 16 : o.end_no_sym;
    return o;
}

// 1 if a tb200_process_decode_host / tb200_decode_batch_host call of this geometry fits the handle's one-shot result region
extern "C" int tb200_host_oneshot_fits(tb200_handle h, int n, int out_stride, int max_frames, int want_symbols)
{
    if (!h || !h->oshot || !h->pin || n <= 0 || out_stride < 0 || max_frames <= 0 || max_frames > 64) return 0;
    const OneShot o = oneshot_layout((size_t)n, (size_t)out_stride, (size_t)max_frames, want_symbols != 0);
    return (o.end <= h->oshot_bytes && o.end <= h->pin_bytes) ? 1 : 0;
}

// process() + decode() of host chunks in ONE call: the class API's per-chunk pattern (ui/modern.py:1919-2077:
// processor.process(samples) then decoder.decode(demodulated)) costs one upload, one kernel sequence with the fused
// tail, ONE download of every result through a pinned mirror and one synchronisation.
extern "C" int tb200_process_decode_host(tb200_handle h, const void *h_iq, int dtype, int n_chunks, int chunk_len,
                                         const double *h_freq_off, uint8_t *h_dibits, void *h_symbols, int out_stride,
                                         int32_t *h_n_dibits, int32_t *h_phase, tb200_frame *h_frames, int max_frames,
                                         int32_t *h_n_frames, int64_t *h_stats, void *stream)
{
    DevGuard dev_guard(h);
    if (!h || !h_n_dibits || !h_frames || !h_n_frames || max_frames <= 0)
        return fail(TB200_E_ARG, "tb200_process_decode_host: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    if (n_chunks <= 0) return TB200_OK;
    if (n_chunks > h->cfg.max_chunks || chunk_len > h->cfg.max_chunk_len)
        return fail(TB200_E_ARG, "tb200_process_decode_host: exceeds the handle capacity");
    const size_t esz = dtype == TB200_DTYPE_C64 ? 8 : (dtype == TB200_DTYPE_SC16 ? 4 : 16);
    const size_t in_bytes = (size_t)n_chunks * chunk_len * esz;
    void *stage = dtype == TB200_DTYPE_SC16 ? h->stage_sc16 : h->stage_in;
    if (in_bytes > (dtype == TB200_DTYPE_SC16 ? h->stage_in_bytes / 2 : h->stage_in_bytes))
        return fail(TB200_E_ARG, "tb200_process_decode_host: input exceeds staging (complex128 needs 2x)");
    const OneShot o = oneshot_layout((size_t)n_chunks, (size_t)out_stride, (size_t)max_frames, h_symbols != nullptr);
    if (!h->oshot || !h->pin || o.end > h->oshot_bytes || o.end > h->pin_bytes || (size_t)out_stride > h->stage_out_stride)
        return fail(TB200_E_ARG, "tb200_process_decode_host: %d chunks x stride %d x %d frames exceed the one-shot result region "
                                 "(this entry point serves the chunk-by-chunk class API; batches use tb200_process_decode_batch)",
                    n_chunks, out_stride, max_frames);
    if (in_bytes) TB_CUDA(cudaMemcpyAsync(stage, h_iq, in_bytes, cudaMemcpyHostToDevice, st));
    double *d_f = nullptr;
    if (h_freq_off) {
        TB_CUDA(cudaMemcpyAsync(h->stage_freq, h_freq_off, sizeof(double) * n_chunks, cudaMemcpyHostToDevice, st));
        d_f = h->stage_freq;
    }
    char *r = h->oshot;
    TB_CUDA(cudaMemsetAsync(r + o.nfr, 0, o.dibits - o.nfr, st));                 // frame counts and stats start at zero
    int rc = tb200_process_decode_batch(h, stage, dtype, n_chunks, chunk_len, chunk_len, d_f, (uint8_t *)(r + o.dibits),
                                        h_symbols ? (void *)(r + o.symbols) : nullptr, out_stride, (int32_t *)(r + o.cnt),
                                        (int32_t *)(r + o.phase), (tb200_frame *)(r + o.frames), max_frames,
                                        (int32_t *)(r + o.nfr), (int64_t *)(r + o.stats), stream);
    if (rc) return rc;
    TB_CUDA(cudaMemcpyAsync(h->pin, r, o.end, cudaMemcpyDeviceToHost, st));
    TB_CUDA(cudaStreamSynchronize(st));
    const char *p = h->pin;
    memcpy(h_n_dibits, p + o.cnt, sizeof(int32_t) * n_chunks);
    if (h_phase) memcpy(h_phase, p + o.phase, sizeof(int32_t) * n_chunks);
    memcpy(h_n_frames, p + o.nfr, sizeof(int32_t) * n_chunks);
    if (h_stats) memcpy(h_stats, p + o.stats, 3 * sizeof(int64_t));
    if (h_dibits) memcpy(h_dibits, p + o.dibits, (size_t)n_chunks * out_stride);
    for (int c = 0; c < n_chunks; ++c) {                                          // only the frames that exist
        int k = ((const int32_t *)(p + o.nfr))[c];
        if (k > max_frames) k = max_frames;
        if (k > 0) memcpy(h_frames + (size_t)c * max_frames, p + o.frames + (size_t)c * max_frames * sizeof(tb200_frame), (size_t)k * sizeof(tb200_frame));
    }
    if (h_symbols) memcpy(h_symbols, p + o.symbols, (size_t)n_chunks * out_stride * 16);
    return TB200_OK;
}

// decode() of host dibit rows in one call (upload, fused decode launch, one download, one synchronisation)
extern "C" int tb200_decode_batch_host(tb200_handle h, const uint8_t *h_dibits, int n_streams, int stride, const int32_t *h_n_dibits,
                                       tb200_frame *h_frames, int max_frames, int32_t *h_n_frames, int64_t *h_stats, void *stream)
{
    DevGuard dev_guard(h);
    if (!h || n_streams < 0 || stride < 0 || !h_n_dibits || !h_frames || !h_n_frames || max_frames <= 0 || (stride && !h_dibits))
        return fail(TB200_E_ARG, "tb200_decode_batch_host: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    if (n_streams == 0) return TB200_OK;
    if (h_stats) h_stats[0] = h_stats[1] = h_stats[2] = 0;
    if (stride == 0) { memset(h_n_frames, 0, sizeof(int32_t) * n_streams); return TB200_OK; }
    const OneShot o = oneshot_layout((size_t)n_streams, 0, (size_t)max_frames, false);
    if ((size_t)n_streams * stride > (size_t)h->max_streams * (h->bits_stride / 2) || n_streams > h->max_streams)
        return fail(TB200_E_ARG, "tb200_decode_batch_host: %d x %d dibits exceed the handle capacity", n_streams, stride);
    if (!h->oshot || !h->pin || o.end > h->oshot_bytes || o.end > h->pin_bytes)
        return fail(TB200_E_ARG, "tb200_decode_batch_host: %d streams x %d frames exceed the one-shot result region", n_streams, max_frames);
    TB_CUDA(cudaMemcpyAsync(h->stage_dec_in, h_dibits, (size_t)n_streams * stride, cudaMemcpyHostToDevice, st));
    TB_CUDA(cudaMemcpyAsync(h->stage_dec_cnt, h_n_dibits, sizeof(int32_t) * n_streams, cudaMemcpyHostToDevice, st));
    char *r = h->oshot;
    TB_CUDA(cudaMemsetAsync(r + o.nfr, 0, o.dibits - o.nfr, st));
    int rc = tb200_decode_batch(h, h->stage_dec_in, n_streams, stride, h->stage_dec_cnt, (tb200_frame *)(r + o.frames), max_frames,
                                (int32_t *)(r + o.nfr), (int64_t *)(r + o.stats), stream);
    if (rc) return rc;
    TB_CUDA(cudaMemcpyAsync(h->pin, r, o.end, cudaMemcpyDeviceToHost, st));
    TB_CUDA(cudaStreamSynchronize(st));
    const char *p = h->pin;
    memcpy(h_n_frames, p + o.nfr, sizeof(int32_t) * n_streams);
    if (h_stats) memcpy(h_stats, p + o.stats, 3 * sizeof(int64_t));
    for (int c = 0; c < n_streams; ++c) {
        int k = h_n_frames[c];
        if (k > max_frames) k = max_frames;
        if (k > 0) memcpy(h_frames + (size_t)c * max_frames, p + o.frames + (size_t)c * max_frames * sizeof(tb200_frame), (size_t)k * sizeof(tb200_frame));
    }
    return TB200_OK;
}

// ------------------------------------------------------------------------- stage level
extern "C" int tb200_decimate(tb200_handle h, const void *d_x, int dtype, int n, int q, void *d_out, void *stream)
{
    DevGuard dev_guard(h);
    if (!h || !d_x || !d_out || q < 2) return fail(TB200_E_ARG, "tb200_decimate: bad argument");
    if (n <= 27) return fail(TB200_E_ARG, "The length of the input vector x must be greater than padlen, which is 27.");
    const size_t esz = dtype == TB200_DTYPE_C64 ? 4 : 8;
    if ((size_t)2 * (n + 54) * esz > h->fwd_bytes) return fail(TB200_E_ARG, "tb200_decimate: n exceeds the handle capacity");
    int rc = get_decim(h, q);
    if (rc) return rc;
    long n_dec = (n + q - 1) / q;
    if (h->cfg.mode == TB200_MODE_FAST && dtype == TB200_DTYPE_C64 && h->decim_tc && tc_decimator_usable(d_x, n, n + (n & 1), q))
        return launch_decimate_tc(h, d_x, 1, n, n + (n & 1), q, h->dec, d_out, n_dec, (cudaStream_t)stream);
    if (h->cfg.mode == TB200_MODE_FAST && dtype == TB200_DTYPE_C64 && fast_decimator_supported(q))
        return launch_decimate_fast(h, d_x, 1, n, n, q, h->dec, d_out, n_dec, (cudaStream_t)stream);
    return launch_decimate_exact(h, d_x, dtype, 1, n, n, q, h->dec, h->fwd, d_out, n_dec, (cudaStream_t)stream);
}

extern "C" int tb200_freq_shift(tb200_handle h, const void *d_x, int dtype, int n, double freq_offset, double fs,
                                void *d_out_c128, void *stream)
{
    DevGuard dev_guard(h);
    if (!h || n < 0 || (n && (!d_x || !d_out_c128))) return fail(TB200_E_ARG, "tb200_freq_shift: bad argument");
    return launch_freq_shift(h, d_x, dtype, n, freq_offset, fs, (double *)d_out_c128, (cudaStream_t)stream);
}

extern "C" int tb200_filter(tb200_handle h, const void *d_x, int dtype, int n, double bandwidth, double fs,
                            void *d_out_c128, void *stream)
{
    DevGuard dev_guard(h);
    if (!h || !d_x || !d_out_c128) return fail(TB200_E_ARG, "tb200_filter: bad argument");
    if (n <= 15) return fail(TB200_E_ARG, "The length of the input vector x must be greater than padlen, which is 15.");
    if (h->cfg.mode == TB200_MODE_FAST && chan_fast_supported(bandwidth, fs))          // time-parallel form: no scratch
        return launch_channel_fast(h, d_x, dtype, 1, n, n, nullptr, fs, bandwidth, (double *)d_out_c128, n, nullptr, nullptr,
                                   0, 1, (cudaStream_t)stream);
    // forward-pass scratch [time][lane]: the channel-filter workspace is sized for the DECIMATED chunk; a full-rate
    // input (filter_signal called directly at the capture rate) borrows the decimator's forward scratch, idle here
    const size_t need = (size_t)2 * (n + 30) * sizeof(double);
    double *scratch = h->chan_fwd;
    if (need > h->chan_bytes) {
        if (need > h->fwd_bytes) return fail(TB200_E_ARG, "tb200_filter: n exceeds the handle capacity");
        scratch = (double *)h->fwd;
    }
    int rc = get_chan(h, bandwidth, fs);
    if (rc) return rc;
    return launch_channel_exact(h, d_x, dtype, 1, n, n, nullptr, fs, h->chan, 1, scratch, (double *)d_out_c128, n,
                                (cudaStream_t)stream);
}

extern "C" int tb200_resample(tb200_handle h, const void *d_x, int dtype, int n, int num, void *d_out, void *stream)
{
    DevGuard dev_guard(h);
    if (!h || n < 0 || num < 0 || (n && num && (!d_x || !d_out))) return fail(TB200_E_ARG, "tb200_resample: bad argument");
    if (dtype != TB200_DTYPE_C64 && dtype != TB200_DTYPE_C128) return fail(TB200_E_ARG, "tb200_resample: dtype");
    // scratch: the decimator workspaces (forward scratch + decimated chunks, carved back to back) are idle here
    if (!h->fwd) return fail(TB200_E_ARG, "tb200_resample: handle has no process workspace");
    size_t avail = (size_t)(((char *)h->dec_out + h->dec_bytes) - (char *)h->fwd);
    return launch_resample(h, d_x, dtype, n, num, d_out, (double *)h->fwd, avail, (cudaStream_t)stream);
}

extern "C" int tb200_extract_symbols(tb200_handle h, const void *d_x_c128, int n, double fs, void *d_out_c128,
                                     int32_t *d_n_out, int32_t *d_phase, void *stream)
{
    DevGuard dev_guard(h);
    if (!h || n < 0 || !d_n_out) return fail(TB200_E_ARG, "tb200_extract_symbols: bad argument");
    int sps = (int)(fs / 18000);
    int step = sps / 8 > 1 ? sps / 8 : 1;
    long cap = sps > 1 ? n / sps + 1 : n;
    return launch_symbols(h, (const double *)d_x_c128, 1, n, n, sps, step, (double *)d_out_c128, nullptr, cap, nullptr,
                          d_n_out, d_phase, (cudaStream_t)stream);
}

extern "C" int tb200_demod_dqpsk(tb200_handle h, const void *d_sym_c128, int n, uint8_t *d_dibits, void *stream)
{
    DevGuard dev_guard(h);
    if (!h || n < 0) return fail(TB200_E_ARG, "tb200_demod_dqpsk: bad argument");
    return launch_slicer(h, (const double *)d_sym_c128, n, d_dibits, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------- decoder
extern "C" int tb200_symbols_to_bits(tb200_handle h, const uint8_t *d_symbols, int n, uint8_t *d_bits, uint8_t *d_mapped,
                                     void *stream)
{
    DevGuard dev_guard(h);
    if (!h || n < 0 || (n && (!d_symbols || !d_bits))) return fail(TB200_E_ARG, "tb200_symbols_to_bits: bad argument");
    if (n == 0) return TB200_OK;
    // mapped is written at bits_stride/2 pitch: with one stream the pitch is irrelevant
    return launch_symbols_to_bits(h, d_symbols, 1, n, nullptr, n, d_bits, d_mapped, 2L * n, (cudaStream_t)stream);
}

extern "C" int tb200_find_sync(tb200_handle h, const uint8_t *d_bits, int n_bits, double threshold, int32_t *d_pos,
                               int max_pos, int32_t *d_count, int32_t *d_max_match, void *stream)
{
    DevGuard dev_guard(h);
    if (!h || n_bits < 0 || !d_count || !d_max_match) return fail(TB200_E_ARG, "tb200_find_sync: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    if (n_bits < 22) {                                               // decoder.py:201-204
        TB_CUDA(cudaMemsetAsync(d_count, 0, sizeof(int32_t), st));
        TB_CUDA(cudaMemsetAsync(d_max_match, 0, sizeof(int32_t), st));
        return TB200_OK;
    }
    if ((long)n_bits > h->bits_stride * h->max_streams) return fail(TB200_E_ARG, "tb200_find_sync: n_bits exceeds the handle capacity");
    int rc = launch_sync_corr(h, d_bits, 1, n_bits, nullptr, n_bits, h->m1, h->m2, st);
    if (rc) return rc;
    return launch_find_sync(h, h->m1, h->m2, n_bits, threshold, d_pos, max_pos, d_count, d_max_match, st);
}

extern "C" int tb200_decode_batch(tb200_handle h, const uint8_t *d_dibits, int n_streams, int stride,
                                  const int32_t *d_n_dibits, tb200_frame *d_frames, int max_frames, int32_t *d_n_frames,
                                  int64_t *d_stats, void *stream)
{
    DevGuard dev_guard(h);
    if (!h || n_streams < 0 || stride < 0 || !d_n_dibits || !d_frames || !d_n_frames)
        return fail(TB200_E_ARG, "tb200_decode_batch: bad argument");
    if (n_streams == 0) return TB200_OK;
    cudaStream_t st = (cudaStream_t)stream;
    long bstride = 2L * stride;
    if ((size_t)n_streams * bstride > (size_t)h->max_streams * h->bits_stride)
        return fail(TB200_E_ARG, "tb200_decode_batch: %d x %d dibits exceed the handle capacity", n_streams, stride);
    if (stride == 0) { TB_CUDA(cudaMemsetAsync(d_n_frames, 0, sizeof(int32_t) * n_streams, st)); return TB200_OK; }
    int rc;
    if (stride <= 6000 && max_frames <= 64) {          // one chunk's worth: fused single-launch path (42 KB smem)
        StageTimer tm(h, TB_STAGE_DECODE, st);         // (keeps its sync positions in shared memory: no sync_pos capacity rule)
        return launch_decode_fused(h, d_dibits, n_streams, stride, d_n_dibits, d_frames, max_frames, d_n_frames, d_stats, st);
    }
    if ((size_t)n_streams * max_frames > (size_t)h->max_streams * h->max_frames)
        return fail(TB200_E_ARG, "tb200_decode_batch: max_frames too large for the handle");
    {
        StageTimer tm(h, TB_STAGE_BITS_SYNC, st);
        rc = launch_symbols_to_bits(h, d_dibits, n_streams, stride, d_n_dibits, 0, h->bits, h->mapped, bstride, st);
        if (rc) return rc;
        rc = launch_sync_corr(h, h->bits, n_streams, bstride, d_n_dibits, 0, h->m1, h->m2, st);
        if (rc) return rc;
    }
    StageTimer tm(h, TB_STAGE_DECODE, st);
    return launch_decode(h, h->bits, h->mapped, h->m1, h->m2, n_streams, bstride, d_n_dibits, d_frames, max_frames,
                         d_n_frames, d_stats, st);
}

extern "C" int tb200_burst_batch(tb200_handle h, const uint8_t *d_dibits, int64_t n, uint8_t *d_rec, uint8_t *d_full,
                                 int64_t *d_stats, void *stream)
{
    DevGuard dev_guard(h);
    if (!h || n < 0 || (n && (!d_dibits || !d_rec))) return fail(TB200_E_ARG, "tb200_burst_batch: bad argument");
    StageTimer tm(h, TB_STAGE_BURST, (cudaStream_t)stream);
    return launch_burst_batch(h, d_dibits, n, d_rec, d_full, d_stats, (cudaStream_t)stream);
}

extern "C" int tb200_burst_batch_host(tb200_handle h, const uint8_t *h_dibits, int64_t n, uint8_t *h_rec, uint8_t *h_full,
                                      int64_t *h_stats, void *stream)
{
    DevGuard dev_guard(h);
    if (!h || n < 0 || (n && (!h_dibits || !h_rec))) return fail(TB200_E_ARG, "tb200_burst_batch_host: bad argument");
    if (n > h->cfg.max_bursts) return fail(TB200_E_ARG, "tb200_burst_batch_host: %lld bursts exceed max_bursts %d", (long long)n, h->cfg.max_bursts);
    if (n == 0) return TB200_OK;
    cudaStream_t st = (cudaStream_t)stream;
    TB_CUDA(cudaMemcpyAsync(h->stage_burst_in, h_dibits, (size_t)n * 255, cudaMemcpyHostToDevice, st));
    if (h_stats) TB_CUDA(cudaMemsetAsync(h->stage_stats, 0, 3 * sizeof(int64_t), st));
    int rc = launch_burst_batch(h, h->stage_burst_in, n, h->stage_burst_rec, h_full ? h->stage_burst_full : nullptr,
                                h_stats ? h->stage_stats : nullptr, st);
    if (rc) return rc;
    TB_CUDA(cudaMemcpyAsync(h_rec, h->stage_burst_rec, (size_t)n * 32, cudaMemcpyDeviceToHost, st));
    if (h_full) TB_CUDA(cudaMemcpyAsync(h_full, h->stage_burst_full, (size_t)n * 64, cudaMemcpyDeviceToHost, st));
    int64_t tmp[3];
    if (h_stats) TB_CUDA(cudaMemcpyAsync(tmp, h->stage_stats, sizeof(tmp), cudaMemcpyDeviceToHost, st));
    TB_CUDA(cudaStreamSynchronize(st));
    if (h_stats) for (int i = 0; i < 3; ++i) h_stats[i] += tmp[i];
    return TB200_OK;
}

extern "C" int tb200_crc16_rows(tb200_handle h, const uint8_t *d_bits, int n_rows, int n_bits, int row_stride,
                                uint16_t *d_crc, void *stream)
{
    DevGuard dev_guard(h);
    if (!h || n_rows < 0 || n_bits < 0 || !d_crc) return fail(TB200_E_ARG, "tb200_crc16_rows: bad argument");
    return launch_crc16_rows(h, d_bits, n_rows, n_bits, row_stride, d_crc, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------- scanner heuristics / codec frames
extern "C" int tb200_detect_batch(tb200_handle h, const void *d_iq, int dtype, int n_dwells, int64_t n, int64_t stride, double fs,
                                  int num_windows, tb200_detect_stats *d_out, void *stream)
{
    DevGuard dev_guard(h);
    if (!h || n_dwells < 0 || n < 0 || stride < n || !(fs > 0) || (n_dwells && (!d_iq || !d_out)) ||
        (dtype != TB200_DTYPE_C64 && dtype != TB200_DTYPE_C128))
        return fail(TB200_E_ARG, "tb200_detect_batch: bad argument");
    // scratch: the decimator's forward scratch is idle here
    return launch_detect(h, d_iq, dtype, n_dwells, n, stride, fs, num_windows, d_out, h->fwd, h->fwd_bytes, (cudaStream_t)stream);
}

extern "C" int tb200_voice_blocks(tb200_handle h, int mode, const void *d_src, int64_t src_stride, const int32_t *d_n_src,
                                  const int32_t *d_frame_stream, const int32_t *d_frame_pos, int64_t n_frames, double scale,
                                  int16_t *d_blocks, uint8_t *d_valid, void *stream)
{
    DevGuard dev_guard(h);
    if (!h || n_frames < 0 || mode < TB200_VOICE_SYMBOLS || mode > TB200_VOICE_SOFT || src_stride < 0 ||
        (n_frames && (!d_src || !d_blocks)) || (n_frames && mode == TB200_VOICE_SYMBOLS && !d_frame_pos))
        return fail(TB200_E_ARG, "tb200_voice_blocks: bad argument");
    return launch_voice_blocks(h, mode, d_src, src_stride, d_n_src, d_frame_stream, d_frame_pos, n_frames, (float)scale, d_blocks,
                               d_valid, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------- etsi mode
extern "C" int tb200_etsi_decode_batch(tb200_handle h, const int8_t *d_soft, int64_t n, int channel, uint32_t scramb_init,
                                       uint8_t *d_bits, uint8_t *d_crc_ok, int32_t *d_metric, void *stream)
{
    DevGuard dev_guard(h);
    if (!h || n < 0 || (n && (!d_soft || !d_bits || !d_crc_ok || !d_metric)))
        return fail(TB200_E_ARG, "tb200_etsi_decode_batch: bad argument");
    StageTimer tm(h, TB_STAGE_ETSI, (cudaStream_t)stream);
    return launch_etsi_decode(h, d_soft, n, channel, scramb_init, d_bits, d_crc_ok, d_metric, (cudaStream_t)stream);
}

extern "C" int tb200_etsi_aach_decode(tb200_handle h, const int8_t *d_soft, int64_t n, const uint32_t *gen_rows, uint32_t scramb_init,
                                      uint16_t *d_bits, int32_t *d_metric, int32_t *d_margin, void *stream)
{
    DevGuard dev_guard(h);
    if (!h || n < 0 || !gen_rows || (n && (!d_soft || !d_bits || !d_metric)))
        return fail(TB200_E_ARG, "tb200_etsi_aach_decode: bad argument");
    StageTimer tm(h, TB_STAGE_ETSI, (cudaStream_t)stream);
    return launch_etsi_aach(h, d_soft, n, gen_rows, scramb_init, d_bits, d_metric, d_margin, (cudaStream_t)stream);
}

extern "C" int tb200_etsi_decode_custom(tb200_handle h, const int8_t *d_soft, int64_t n, int k1, int k2, int k3, int a, int rate,
                                        int crc, uint32_t scramb_init, uint8_t *d_bits, uint8_t *d_crc_ok, int32_t *d_metric,
                                        void *stream)
{
    DevGuard dev_guard(h);
    if (!h || n < 0 || (n && (!d_soft || !d_bits || !d_crc_ok || !d_metric)))
        return fail(TB200_E_ARG, "tb200_etsi_decode_custom: bad argument");
    StageTimer tm(h, TB_STAGE_ETSI, (cudaStream_t)stream);
    return launch_etsi_decode_custom(h, d_soft, n, k1, k2, k3, a, rate, crc, scramb_init, d_bits, d_crc_ok, d_metric,
                                     (cudaStream_t)stream);
}

extern "C" int tb200_etsi_demod_batch(tb200_handle h, const void *d_iq_c64, int n_streams, int64_t n_samples, int64_t stride,
                                      double fs, void *d_sym_c64, float *d_soft, int64_t soft_stride, int32_t *d_n_soft,
                                      void *stream)
{
    DevGuard dev_guard(h);
    if (!h || n_streams < 0 || n_samples < 0 || !d_iq_c64 || !d_soft || !d_n_soft || soft_stride < 2)
        return fail(TB200_E_ARG, "tb200_etsi_demod_batch: bad argument");
    StageTimer tm(h, TB_STAGE_ETSI, (cudaStream_t)stream);
    return launch_etsi_demod(h, (const float2 *)d_iq_c64, n_streams, n_samples, stride, fs, h->etsi_taps, (float2 *)d_sym_c64,
                             d_soft, soft_stride, d_n_soft, nullptr, 0, (cudaStream_t)stream);
}

extern "C" int64_t tb200_etsi_demod_scratch_bytes(int n_streams, int64_t n_samples, double fs, int want_symbols)
{
    if (n_streams <= 0 || n_samples <= 0 || !(fs > 0.0)) return 0;
    return etsi_demod_scratch_bytes(n_streams, n_samples, fs, want_symbols);
}

extern "C" int tb200_etsi_demod_segmented(tb200_handle h, const void *d_iq_c64, int n_streams, int64_t n_samples, int64_t stride,
                                          double fs, void *d_sym_c64, float *d_soft, int64_t soft_stride, int32_t *d_n_soft,
                                          void *d_scratch, int64_t scratch_bytes, void *stream)
{
    DevGuard dev_guard(h);
    if (!h || n_streams < 0 || n_samples < 0 || !d_iq_c64 || !d_soft || !d_n_soft || soft_stride < 2 || scratch_bytes < 0)
        return fail(TB200_E_ARG, "tb200_etsi_demod_segmented: bad argument");
    StageTimer tm(h, TB_STAGE_ETSI, (cudaStream_t)stream);
    return launch_etsi_demod(h, (const float2 *)d_iq_c64, n_streams, n_samples, stride, fs, h->etsi_taps, (float2 *)d_sym_c64,
                             d_soft, soft_stride, d_n_soft, d_scratch, (size_t)scratch_bytes, (cudaStream_t)stream);
}

extern "C" int tb200_etsi_find_bursts(tb200_handle h, const float *d_soft, int n_streams, int64_t soft_stride,
                                      const int32_t *d_n_soft, double threshold, float *d_metric, int8_t *d_blocks,
                                      int32_t *d_burst_pos, int32_t *d_n_bursts, int max_bursts, void *stream)
{
    DevGuard dev_guard(h);
    if (!h || n_streams < 0 || !d_soft || !d_n_soft || !d_metric || !d_blocks || !d_burst_pos || !d_n_bursts || max_bursts <= 0)
        return fail(TB200_E_ARG, "tb200_etsi_find_bursts: bad argument");
    StageTimer tm(h, TB_STAGE_ETSI, (cudaStream_t)stream);
    return launch_etsi_find_bursts(h, d_soft, n_streams, soft_stride, d_n_soft, (float)threshold, d_metric, d_blocks, d_burst_pos,
                                   d_n_bursts, max_bursts, (cudaStream_t)stream);
}

extern "C" int tb200_etsi_find_sequence(tb200_handle h, const float *d_soft, int n_streams, int64_t soft_stride,
                                        const int32_t *d_n_soft, int sequence, double threshold, int min_gap_symbols,
                                        float *d_metric, int32_t *d_pos, float *d_corr, int32_t *d_n_hits, int max_hits,
                                        void *stream)
{
    DevGuard dev_guard(h);
    if (!h || n_streams < 0 || !d_soft || !d_n_soft || !d_metric || !d_pos || !d_corr || !d_n_hits || max_hits <= 0 ||
        sequence < 0 || sequence > 4 || soft_stride < 2)
        return fail(TB200_E_ARG, "tb200_etsi_find_sequence: bad argument");
    StageTimer tm(h, TB_STAGE_ETSI, (cudaStream_t)stream);
    return launch_etsi_find_sequence(h, d_soft, n_streams, soft_stride, d_n_soft, sequence, (float)threshold, min_gap_symbols,
                                     d_metric, d_pos, d_corr, d_n_hits, max_hits, (cudaStream_t)stream);
}

extern "C" int tb200_etsi_extract_bits(tb200_handle h, const float *d_soft, int n_streams, int64_t soft_stride,
                                       const int32_t *d_n_soft, const int32_t *d_burst_pos, const int32_t *d_n_bursts,
                                       int max_bursts, int off_a, int len_a, int off_b, int len_b, int8_t *d_out, void *stream)
{
    DevGuard dev_guard(h);
    if (!h || n_streams < 0 || !d_soft || !d_n_soft || !d_burst_pos || !d_n_bursts || !d_out || max_bursts <= 0 || len_a < 0 ||
        len_b < 0 || len_a + len_b <= 0 || off_a < 0 || off_b < 0 || off_a + len_a > 510 || off_b + len_b > 510)
        return fail(TB200_E_ARG, "tb200_etsi_extract_bits: bad argument");
    StageTimer tm(h, TB_STAGE_ETSI, (cudaStream_t)stream);
    return launch_etsi_extract(h, d_soft, n_streams, soft_stride, d_n_soft, d_burst_pos, d_n_bursts, max_bursts, off_a, len_a,
                               off_b, len_b, d_out, (cudaStream_t)stream);
}

extern "C" int tb200_etsi_afc(tb200_handle h, float *d_soft, int n_streams, int64_t soft_stride, const int32_t *d_n_soft,
                              float *d_theta, void *stream)
{
    DevGuard dev_guard(h);
    if (!h || n_streams < 0 || !d_soft || !d_n_soft || (soft_stride & 1) || (reinterpret_cast<uintptr_t>(d_soft) & 7))
        return fail(TB200_E_ARG, "tb200_etsi_afc: bad argument");
    StageTimer tm(h, TB_STAGE_ETSI, (cudaStream_t)stream);
    return launch_etsi_afc(h, d_soft, n_streams, soft_stride, d_n_soft, d_theta, (cudaStream_t)stream);
}

extern "C" int tb200_etsi_find_sync_bursts(tb200_handle h, const float *d_soft, int n_streams, int64_t soft_stride,
                                           const int32_t *d_n_soft, double threshold, float *d_metric, int8_t *d_sb1,
                                           int8_t *d_bkn2, int32_t *d_burst_pos, int32_t *d_n_bursts, int max_bursts, void *stream)
{
    DevGuard dev_guard(h);
    if (!h || n_streams < 0 || !d_soft || !d_n_soft || !d_metric || !d_sb1 || !d_burst_pos || !d_n_bursts || max_bursts <= 0)
        return fail(TB200_E_ARG, "tb200_etsi_find_sync_bursts: bad argument");
    StageTimer tm(h, TB_STAGE_ETSI, (cudaStream_t)stream);
    return launch_etsi_find_sync(h, d_soft, n_streams, soft_stride, d_n_soft, (float)threshold, d_metric, d_sb1, d_bkn2, d_burst_pos,
                                 d_n_bursts, max_bursts, (cudaStream_t)stream);
}

extern "C" int tb200_channelize(tb200_handle h, const void *d_iq_c64, int64_t n_samples, int n_channels, void *d_out_c64,
                                int64_t out_stride, int64_t n_frames, void *stream)
{
    DevGuard dev_guard(h);
    if (!h || !d_iq_c64 || !d_out_c64 || n_samples < 0 || n_channels < 2 || (n_channels & 1) || n_channels > h->cfg.max_channels)
        return fail(TB200_E_ARG, "tb200_channelize: bad argument (n_channels must be even and <= the handle's max_channels)");
    if (n_frames < 0 || n_frames > out_stride) return fail(TB200_E_ARG, "tb200_channelize: n_frames exceeds out_stride");
    StageTimer tm(h, TB_STAGE_ETSI, (cudaStream_t)stream);
    return launch_channelizer(h, (const float2 *)d_iq_c64, n_samples, n_channels, h->chan_proto, (float2 *)h->chan_tw,
                              (float2 *)d_out_c64, out_stride, n_frames, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------- reference-mode channel bank
extern "C" int tb200_refchan_batch(tb200_handle h, const void *d_iq_c64, int64_t n_samples, int n_channels, int q, double fs,
                                   void *d_out_c64, int64_t out_stride, void *stream)
{
    DevGuard dev_guard(h);
    if (!h || !d_iq_c64 || !d_out_c64 || n_samples <= 0 || n_channels < 2 || q < 2 || !(fs > 0))
        return fail(TB200_E_ARG, "tb200_refchan_batch: bad argument");
    int rc = get_decim(h, q);
    if (rc) return rc;
    StageTimer tm(h, TB_STAGE_DECIM, (cudaStream_t)stream);
    return launch_refchan(h, (const float2 *)d_iq_c64, (long)n_samples, n_channels, q, fs, h->dec, (float2 *)d_out_c64, (long)out_stride,
                          (cudaStream_t)stream);
}

// ------------------------------------------------------------------------- PSD
extern "C" int tb200_psd_batch(tb200_handle h, const void *d_iq_c64, int64_t n_frames, int64_t frame_stride, int n_fft,
                               double sample_rate, double ema_alpha, int first, float *d_power, float *d_smoothed,
                               double *d_ema_state, double *d_stats, void *stream)
{
    DevGuard dev_guard(h);
    if (!h || n_frames < 0 || !d_iq_c64) return fail(TB200_E_ARG, "tb200_psd_batch: bad argument");
    if (n_fft < 16 || (n_fft & (n_fft - 1)) || n_fft > h->cfg.max_fft)
        return fail(TB200_E_ARG, "tb200_psd_batch: n_fft must be a power of two in [16, %d]", h->cfg.max_fft);
    if (d_smoothed && (!d_power || !d_ema_state)) return fail(TB200_E_ARG, "tb200_psd_batch: smoothing needs d_power and d_ema_state");
    StageTimer tm(h, TB_STAGE_PSD, (cudaStream_t)stream);
    return launch_psd(h, (const float *)d_iq_c64, n_frames, frame_stride, n_fft, sample_rate, ema_alpha, first, d_power,
                      d_smoothed, d_ema_state, d_stats, h->psd_tables, (cudaStream_t)stream, h->psd_f32, h->psd_avg > 1 ? h->psd_avg : 1);
}

// ------------------------------------------------------------------------- debugging aid
extern "C" int tb200_debug_read(tb200_handle h, int region, void *h_dst, size_t nbytes)
{
    DevGuard dev_guard(h);
    if (!h || !h_dst) return fail(TB200_E_ARG, "tb200_debug_read: bad argument");
    const void *src = nullptr;
    size_t cap = 0;
    switch (region) {
    case 0: src = h->dec_out; cap = h->dec_bytes; break;
    case 1: src = h->filtered; cap = h->filt_bytes; break;
    case 2: src = h->fwd; cap = h->fwd_bytes; break;
    case 3: src = h->chan_fwd; cap = h->chan_bytes; break;
    case 4: src = h->etsi_taps; cap = 32 * 32 * sizeof(float); break;
    default: return fail(TB200_E_ARG, "tb200_debug_read: unknown region %d", region);
    }
    if (nbytes > cap) return fail(TB200_E_ARG, "tb200_debug_read: %zu bytes > region size %zu", nbytes, cap);
    TB_CUDA(cudaDeviceSynchronize());
    TB_CUDA(cudaMemcpy(h_dst, src, nbytes, cudaMemcpyDeviceToHost));
    return TB200_OK;
}
