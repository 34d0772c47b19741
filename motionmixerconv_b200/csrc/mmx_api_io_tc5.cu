// Embedding (per-frame Linear) and output head of the MlpMixer on the Blackwell tensor cores: launch layer of
// mmx_lin_tc5.cuh and mmx_head_tc5.cuh.  Reached from mmx_linear_{fwd,bwd}_prec / mmx_mlp_head_{fwd,bwd}_prec when the caller
// asks for the reduced-precision mode (MMX_PREC_TF32) and the shape is served; everything else stays on the fp32 kernels.
#include "mmx_launch.cuh"

#if defined(MMX_HOST_EMU)
bool mmx_lin_tc5_ok(long long, int, int, const void*, const void*, const void*) { return false; }
int mmx_lin_tc5_fwd(long long, int, int, const float*, const float*, const float*, float*, void*) { return fail(MMX_E_UNSUPPORTED, "no tensor cores in the emulator"); }
int mmx_lin_tc5_bwd(long long, int, int, const float*, const float*, const float*, float*, float*, float*, void*) { return fail(MMX_E_UNSUPPORTED, "no tensor cores in the emulator"); }
bool mmx_head_tc5_ok(const MmxMlpHeadDesc*) { return false; }
int mmx_head_tc5_fwd(const MmxMlpHeadDesc*, const MmxMlpHeadParams*, const float*, float*, void*) { return fail(MMX_E_UNSUPPORTED, "no tensor cores in the emulator"); }
int mmx_head_tc5_bwd(const MmxMlpHeadDesc*, const MmxMlpHeadParams*, const MmxMlpHeadParams*, const float*, const float*, float*, void*) {
    return fail(MMX_E_UNSUPPORTED, "no tensor cores in the emulator");
}
#else
#include "mmx_head_tc5.cuh"
#include "mmx_lin_tc5.cuh"
#include "mmx_tc5_launch.cuh"

using namespace mmx;

// ------------------------------------------------------------------------------------------ linear layer on rows
bool mmx_lin_tc5_ok(long long rows, int K, int N, const void* x, const void* y, const void* dx) {
    if (env_int("MMX_IO_NO_TC5", 0)) return false;
    if (rows <= 0 || K < 8 || N < 8 || (K & 1) || (N & 1)) return false;
    if (K + 1 > 80 || N > 80) return false;                          // one padded operand width serves both sides (+ the ones column)
    if (((rows * K) & 3) || ((rows * N) & 3)) return false;          // bulk copies: every tile is a multiple of 16 bytes
    if ((((uintptr_t)x) | ((uintptr_t)y) | ((uintptr_t)dx)) & 15) return false;
    return true;
}

static void fill_lin(lin::LinArgs& a, long long rows, int K, int N) {
    a.R = rows; a.K = K; a.N = N;
    a.abort_count = mmx_tc5_abort_ptr();
    a.dy = nullptr; a.g_w = a.g_b = nullptr;
}

int mmx_lin_tc5_fwd(long long rows, int K, int N, const float* x, const float* w, const float* b, float* y, void* stream) {
    lin::LinArgs a;
    fill_lin(a, rows, K, N);
    a.x = x; a.w = w; a.b = b; a.out = y;
    const DevInfo di = dev_info();
    const long long ntiles = (rows + 127) / 128;
    const bool small = (K + 1 > N ? K + 1 : N) <= 64;
    const size_t smem = small ? lin::lin_smem_bytes<64>(K, N, false) : lin::lin_smem_bytes<80>(K, N, false);
    const int per_sm = imax(1, imin(2, (int)((di.max_smem + 1024) / (smem + 1024))));
    const int grid = (int)(ntiles < (long long)di.sms * per_sm ? ntiles : (long long)di.sms * per_sm);
    return small ? launch_tc5(lin::lin_fwd_kernel<64, 2>, a, grid, chan::kThreadsChan, smem, stream)
                 : launch_tc5(lin::lin_fwd_kernel<80, 2>, a, grid, chan::kThreadsChan, smem, stream);
}

int mmx_lin_tc5_bwd(long long rows, int K, int N, const float* x, const float* w, const float* dy, float* dw, float* db, float* dx,
                    void* stream) {
    lin::LinArgs a;
    fill_lin(a, rows, K, N);
    a.x = x; a.w = w; a.b = nullptr; a.dy = dy; a.out = dx; a.g_w = dw; a.g_b = db;
    const DevInfo di = dev_info();
    const long long ntiles = (rows + 127) / 128;
    const bool small = (K + 1 > N ? K + 1 : N) <= 64;
    const size_t smem = small ? lin::lin_smem_bytes<64>(K, N, true) : lin::lin_smem_bytes<80>(K, N, true);
    const int per_sm = imax(1, imin(2, (int)((di.max_smem + 1024) / (smem + 1024))));
    const int grid = (int)(ntiles < (long long)di.sms * per_sm ? ntiles : (long long)di.sms * per_sm);
    return small ? launch_tc5(lin::lin_bwd_kernel<64, 2>, a, grid, chan::kThreadsChan, smem, stream)
                 : launch_tc5(lin::lin_bwd_kernel<80, 2>, a, grid, chan::kThreadsChan, smem, stream);
}

// ------------------------------------------------------------------------------------------ output head
bool mmx_head_tc5_ok(const MmxMlpHeadDesc* d) {
    if (env_int("MMX_IO_NO_TC5", 0)) return false;
    if (d->B <= 0 || d->T < 1 || d->To < 1 || d->T > 128 || d->To > 128) return false;
    if (d->H < 8 || d->D < 8 || (d->H & 1) || (d->D & 1) || d->H + 1 > head::KH || d->D > head::KD) return false;
    if (((d->T * d->H) & 3) || ((d->To * d->D) & 3)) return false;
    return true;
}

static void fill_head(head::HeadArgs& a, const MmxMlpHeadDesc* d, const MmxMlpHeadParams* w, const MmxMlpHeadParams* g) {
    a.ln_g = w->ln_w; a.ln_b = w->ln_b; a.wt = w->wt; a.bt = w->bt; a.wf = w->wf; a.bf = w->bf;
    if (g) { a.g_ln_g = g->ln_w; a.g_ln_b = g->ln_b; a.g_wt = g->wt; a.g_bt = g->bt; a.g_wf = g->wf; a.g_bf = g->bf; }
    else a.g_ln_g = a.g_ln_b = a.g_wt = a.g_bt = a.g_wf = a.g_bf = nullptr;
    a.B = d->B; a.T = d->T; a.To = d->To; a.H = d->H; a.D = d->D;
    a.S = 128 / imax(d->T, d->To);
    a.dout = nullptr;
    a.abort_count = mmx_tc5_abort_ptr();
}

int mmx_head_tc5_fwd(const MmxMlpHeadDesc* d, const MmxMlpHeadParams* w, const float* x, float* out, void* stream) {
    head::HeadArgs a;
    fill_head(a, d, w, nullptr);
    a.x = x; a.out = out;
    const DevInfo di = dev_info();
    const int ntiles = (d->B + a.S - 1) / a.S;
    const size_t smem = head::head_smem(d->H, d->D, false).total;
    return launch_tc5(head::head_fwd_kernel, a, imin(ntiles, di.sms), chan::kThreadsChan, smem, stream);
}

int mmx_head_tc5_bwd(const MmxMlpHeadDesc* d, const MmxMlpHeadParams* w, const MmxMlpHeadParams* grads, const float* x, const float* dout,
                     float* dx, void* stream) {
    head::HeadArgs a;
    fill_head(a, d, w, grads);
    a.x = x; a.dout = dout; a.out = dx;
    const DevInfo di = dev_info();
    const int ntiles = (d->B + a.S - 1) / a.S;
    const size_t smem = head::head_smem(d->H, d->D, true).total;
    return launch_tc5(head::head_bwd_kernel, a, imin(ntiles, di.sms), chan::kThreadsChan, smem, stream);
}
#endif
