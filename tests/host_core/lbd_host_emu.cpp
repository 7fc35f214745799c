/* tests/host_core/lbd_host_emu.cpp -- cube_slam_b200/csrc/cs_lbd.cu ITSELF (host orchestration + kernels), compiled for the host:
 * the CUDA execution model and the handful of runtime calls are emulated (cuda_emu.h), and the three detector entry points cs_lbd.cu
 * calls in other translation units (cs_lsd_run_host, cs_edl_run_keylines, cs_edl_sobel_maps) are stand-ins that answer from the CPU oracle
 * in exactly the layout the real ones leave in HBM (cap-strided segment rows, counts, {direction, numOfPixels bits} pairs, int16 maps).
 * tests/test_lbd_host_emu.py then drives the library's real entry points -- cs_detect_descrip_lines_batch, cs_lbd_compute_batch,
 * cs_match_line_descrip_batch -- through the Python mirror and runs the GPU parity test's own assertions on them.  What remains untested
 * without a GPU after this: the real CUDA runtime's behaviour and the detector-side kernels' two new stores.  Test infrastructure, never
 * shipped.  g++ -std=c++20 -O2 -ffp-contract=off -pthread -x c++, linked against oracle/_build/liboracle.so. */
#include "cuda_emu.h"

#include <cstdarg>
#include <cstdio>
#include <string>

#include "../../include/cube_slam_b200.h"

struct cs_ctx {
    void *lbd = nullptr;
    char err[512] = {0};
    int64_t launches = 0;
    /* what the detector stand-ins hand out */
    std::vector<float> lines, extra;
    std::vector<int32_t> counts;
    std::vector<int16_t> dx, dy;
};
cudaStream_t cs_ctx_stream(cs_ctx *) { return nullptr; }

#include "../../cube_slam_b200/csrc/cs_internal.h"

int cs_ctx_device(cs_ctx *) { return 0; }
int cs_ctx_fail(cs_ctx *c, int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(c->err, sizeof c->err, fmt, ap);
    va_end(ap);
    return code;
}
void **cs_ctx_lbd_slot(cs_ctx *c) { return &c->lbd; }
void cs_ctx_count_launches(cs_ctx *c, int64_t n) { c->launches += n; }

extern "C" int lsd_orc_detect(const uint8_t *img, int w, int h, int stride, int channels, float line_length_thres, float *lines_out, int cap, float *raw_lines,
                              int cap_raw, int *n_raw_out, double *scaled_out, double *modgrad_out, double *angles_out, int32_t *list_out, int *list_len,
                              int refine_mode);
extern "C" int edl_orc_detect_keylines(const uint8_t *img, int w, int h, int stride, int channels, float line_length_thres, float *lines_out, float *kl_out,
                                       int cap, int16_t *dx_out, int16_t *dy_out);

static void sobel_of(cs_ctx *c, const uint8_t *imgs, int n_frames, int w, int h, int stride, int channels)
{
    c->dx.assign((size_t)n_frames * w * h, 0);
    c->dy.assign((size_t)n_frames * w * h, 0);
    std::vector<float> l(4 * 16), k(3 * 16);
    for (int f = 0; f < n_frames; f++)
        edl_orc_detect_keylines(imgs + (size_t)f * h * stride, w, h, stride, channels, 1e9f, l.data(), k.data(), 16, &c->dx[(size_t)f * w * h], &c->dy[(size_t)f * w * h]);
}

int cs_lsd_run_host(cs_ctx *c, const uint8_t *imgs, int n_frames, int w, int h, int stride, int channels, float line_length_thres, int cap,
                    const float **d_lines, const int32_t **d_counts, const uint8_t **d_frames)
{
    c->lines.assign((size_t)n_frames * cap * 4, -777.f); /* slots past the count hold garbage on the device too */
    c->counts.assign((size_t)n_frames, 0);
    for (int f = 0; f < n_frames; f++)
        c->counts[f] = lsd_orc_detect(imgs + (size_t)f * h * stride, w, h, stride, channels, line_length_thres, &c->lines[(size_t)f * cap * 4], cap, nullptr, 0, nullptr,
                                      nullptr, nullptr, nullptr, nullptr, nullptr, 2);
    *d_lines = c->lines.data();
    *d_counts = c->counts.data();
    *d_frames = imgs;
    return CS_OK;
}

int cs_edl_run_keylines(cs_ctx *c, const uint8_t *imgs, bool, int n_frames, int w, int h, int stride, int channels, float line_length_thres, int cap,
                        const float **d_lines, const int32_t **d_counts, const float **d_extra, const int16_t **d_dx, const int16_t **d_dy)
{
    c->lines.assign((size_t)n_frames * cap * 4, -777.f);
    c->extra.assign((size_t)n_frames * cap * 2, -777.f);
    c->counts.assign((size_t)n_frames, 0);
    c->dx.assign((size_t)n_frames * w * h, 0);
    c->dy.assign((size_t)n_frames * w * h, 0);
    std::vector<float> kl((size_t)cap * 3);
    for (int f = 0; f < n_frames; f++) {
        const int n = edl_orc_detect_keylines(imgs + (size_t)f * h * stride, w, h, stride, channels, line_length_thres, &c->lines[(size_t)f * cap * 4], kl.data(), cap,
                                              &c->dx[(size_t)f * w * h], &c->dy[(size_t)f * w * h]);
        c->counts[f] = n;
        for (int k = 0; k < n && k < cap; k++) {
            c->extra[((size_t)f * cap + k) * 2] = kl[3 * k];
            const int32_t npx = (int32_t)kl[3 * k + 2];
            std::memcpy(&c->extra[((size_t)f * cap + k) * 2 + 1], &npx, 4);
        }
    }
    *d_lines = c->lines.data();
    *d_counts = c->counts.data();
    *d_extra = c->extra.data();
    *d_dx = c->dx.data();
    *d_dy = c->dy.data();
    return CS_OK;
}

int cs_edl_sobel_maps(cs_ctx *c, const uint8_t *imgs, bool, int n_frames, int w, int h, int stride, int channels, const int16_t **d_dx, const int16_t **d_dy)
{
    sobel_of(c, imgs, n_frames, w, h, stride, channels);
    *d_dx = c->dx.data();
    *d_dy = c->dy.data();
    return CS_OK;
}

#include "../../cube_slam_b200/csrc/cs_lbd.cu"

extern "C" cs_ctx *emu_ctx_new() { return new cs_ctx(); }
extern "C" void emu_ctx_free(cs_ctx *c)
{
    if (c->lbd) cs_lbd_destroy(c->lbd);
    delete c;
}
extern "C" const char *emu_last_error(cs_ctx *c) { return c->err; }
extern "C" long long emu_launches(cs_ctx *c) { return c->launches; }

#ifdef CS_EMU_WITH_SHIM_GLUE
/* The rest of the C ABI that shim/line_lbd_b200.cpp binds, so that the C++ shim and its driver (shim/test/line_shim_driver.cpp) can be
 * linked against this emulated build and run on the CPU (tests/test_lbd_host_emu.py): context management as trivial stand-ins, cs_detect_lines
 * answered from the oracle like the detector entry points above. */
extern "C" int edl_orc_detect(const uint8_t *img, int w, int h, int stride, int channels, float line_length_thres, float *lines_out, int cap, float *raw_lines,
                              int cap_raw, int *n_raw_out, uint8_t *blur_out, int16_t *dx_out, int16_t *dy_out, int16_t *g_out, uint8_t *dir_out,
                              int32_t *anchors_out, int *n_anchors_out, uint8_t *edge_out);
extern "C" cs_ctx *cs_create(int, int, int, int, int, int) { return new cs_ctx(); }
extern "C" void cs_destroy(cs_ctx *c) { emu_ctx_free(c); }
extern "C" const char *cs_last_error(const cs_ctx *c) { return c ? c->err : "null context"; }
extern "C" void cs_default_line_params(cs_line_params *p)
{
    p->use_LSD = 0;
    p->numoctaves = 1;
    p->octaveratio = 1.f;
    p->line_length_thres = 50;
}
extern "C" int cs_detect_lines(cs_ctx *c, const uint8_t *img, int width, int height, int stride, int channels, const cs_line_params *params, float *lines_xyxy,
                               int32_t *n_inout)
{
    const int cap = *n_inout;
    const int n = params->use_LSD ? lsd_orc_detect(img, width, height, stride, channels, params->line_length_thres, lines_xyxy, cap, nullptr, 0, nullptr, nullptr,
                                                   nullptr, nullptr, nullptr, nullptr, 2)
                                  : edl_orc_detect(img, width, height, stride, channels, params->line_length_thres, lines_xyxy, cap, nullptr, 0, nullptr, nullptr,
                                                   nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
    if (n < 0 || n > cap) return cs_ctx_fail(c, CS_ERR_CAPACITY, "line detection failed (%d)", n);
    *n_inout = n;
    return CS_OK;
}
#endif
