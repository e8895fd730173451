// C-ABI entry points of the convolution family: argument validation and the choice between the tcgen05 kernels
// and the CUDA-core kernels (B200_IMPL_AUTO picks tcgen05 whenever the shape allows it).
#include <atomic>
#include <cstdio>
#include <cstdlib>

#include "conv_impl.h"

using namespace b200;

namespace {

int check_conv_fwd(const b200_conv_fwd_params* p) {
  B200_REQUIRE(p, "conv_fwd: null params");
  B200_REQUIRE(p->num_src >= 1 && p->num_src <= 2, "conv_fwd: num_src must be 1 or 2");
  B200_REQUIRE(p->taps == 9 || p->taps == 1, "conv_fwd: taps must be 9 or 1");
  B200_REQUIRE(p->pad == 0 || (p->pad == 1 && p->taps == 9), "conv_fwd: pad must be 0, or 1 for a 3x3");
  B200_REQUIRE(view_ok(&p->dst), "conv_fwd: bad dst view");
  const int k = p->taps == 9 ? 3 : 1;
  for (int i = 0; i < p->num_src; ++i) {
    B200_REQUIRE(view_ok(&p->src[i]), "conv_fwd: bad src[%d] view", i);
    B200_REQUIRE((p->src[i].lo != nullptr) == (p->dst.lo != nullptr),
                 "conv_fwd: sources and destination must be of the same precision tier (lo planes)");
    B200_REQUIRE(p->src[i].n == p->dst.n && p->src[i].h + 2 * p->pad - (k - 1) == p->dst.h &&
                     p->src[i].w + 2 * p->pad - (k - 1) == p->dst.w,
                 "conv_fwd: src[%d] extent %dx%d does not produce dst %dx%d", i, p->src[i].h, p->src[i].w, p->dst.h,
                 p->dst.w);
  }
  return 0;
}

int check_conv_dgrad(const b200_conv_dgrad_params* p) {
  B200_REQUIRE(p, "conv_dgrad: null params");
  B200_REQUIRE(p->num_dst >= 1 && p->num_dst <= 2, "conv_dgrad: num_dst must be 1 or 2");
  B200_REQUIRE(p->taps == 9 || p->taps == 1, "conv_dgrad: taps must be 9 or 1");
  B200_REQUIRE(p->pad == 0 || (p->pad == 1 && p->taps == 9), "conv_dgrad: pad must be 0, or 1 for a 3x3");
  B200_REQUIRE(view_ok(&p->dz), "conv_dgrad: bad dz view");
  const int k = p->taps == 9 ? 3 : 1;
  for (int i = 0; i < p->num_dst; ++i) {
    B200_REQUIRE(view_ok(&p->dst[i]), "conv_dgrad: bad dst[%d] view", i);
    B200_REQUIRE(p->dst[i].n == p->dz.n && p->dst[i].h + 2 * p->pad - (k - 1) == p->dz.h &&
                     p->dst[i].w + 2 * p->pad - (k - 1) == p->dz.w,
                 "conv_dgrad: dst[%d] extent does not match dz", i);
  }
  return 0;
}

int check_conv_wgrad(const b200_conv_wgrad_params* p) {
  B200_REQUIRE(p, "conv_wgrad: null params");
  B200_REQUIRE(p->num_src >= 1 && p->num_src <= 2, "conv_wgrad: num_src must be 1 or 2");
  B200_REQUIRE(p->taps == 9 || p->taps == 1, "conv_wgrad: taps must be 9 or 1");
  B200_REQUIRE(view_ok(&p->dz) && p->dw_f32, "conv_wgrad: bad dz / dw");
  const int k = p->taps == 9 ? 3 : 1;
  for (int i = 0; i < p->num_src; ++i) {
    B200_REQUIRE(view_ok(&p->src[i]), "conv_wgrad: bad src[%d] view", i);
    B200_REQUIRE(p->src[i].n == p->dz.n && p->src[i].h + 2 * p->pad - (k - 1) == p->dz.h &&
                     p->src[i].w + 2 * p->pad - (k - 1) == p->dz.w,
                 "conv_wgrad: src[%d] extent does not match dz", i);
  }
  return 0;
}

int cin_of(const b200_conv_wgrad_params* p) {
  int c = 0;
  for (int i = 0; i < p->num_src; ++i) c += p->src[i].c;
  return c;
}

// B200_IMPL_AUTO silently dropping from the tcgen05 kernels to the CUDA-core `direct` ones is a 10-50x performance cliff
// (channel counts that are not multiples of 8, misaligned views).  Small channel counts (<= 7: the image side of the
// first layer) are served by design; anything wider is counted and reported once per entry point on stderr
// (B200UNET_QUIET=1 silences it; b200unet_fallback_count() lets a host binding surface it).
std::atomic<unsigned long long> g_fallbacks{0};
void note_fallback(const char* what, int channels) {
  if (channels <= 7) return;
  g_fallbacks.fetch_add(1);
  static std::atomic<unsigned> warned{0};
  static const bool quiet = getenv("B200UNET_QUIET") != nullptr;
  unsigned bit = 1u;
  for (const char* c = what; *c; ++c) bit = bit * 31u + (unsigned)*c;
  bit = 1u << (bit % 32u);
  if (!quiet && !(warned.fetch_or(bit) & bit))
    fprintf(stderr, "[b200unet] %s with %d channels does not qualify for the tcgen05 kernel (channel counts must be multiples of 8, "
            "views 16-byte aligned): running the CUDA-core fallback, expect 10-50x lower throughput.  Pad the channels "
            "(b200unet.UNet does this itself) or set B200UNET_QUIET=1.\n", what, channels);
}

template <class P, class OkFn>
int resolve(const P* p, OkFn ok, const char* what, int* impl) {
  *impl = p->impl;
  if (p->impl == B200_IMPL_AUTO) *impl = ok(p) ? B200_IMPL_UMMA : B200_IMPL_DIRECT;
  if (*impl == B200_IMPL_UMMA && !ok(p)) return fail(-1, "%s: shape not supported by the tcgen05 kernel", what);
  B200_REQUIRE(*impl == B200_IMPL_UMMA || *impl == B200_IMPL_DIRECT, "%s: bad impl selector", what);
  return 0;
}

}  // namespace

extern "C" {

unsigned long long b200unet_fallback_count(void) { return g_fallbacks.load(); }

int b200unet_conv_fwd_impl(const b200_conv_fwd_params* p) {
  if (check_conv_fwd(p)) return -1;
  return umma_conv_fwd_ok(p) ? B200_IMPL_UMMA : B200_IMPL_DIRECT;
}
int b200unet_conv_dgrad_impl(const b200_conv_dgrad_params* p) {
  if (check_conv_dgrad(p)) return -1;
  return umma_conv_dgrad_ok(p) ? B200_IMPL_UMMA : B200_IMPL_DIRECT;
}
int b200unet_conv_wgrad_impl(const b200_conv_wgrad_params* p) {
  if (check_conv_wgrad(p)) return -1;
  return umma_conv_wgrad_ok(p) ? B200_IMPL_UMMA : B200_IMPL_DIRECT;
}

int b200unet_convt_fwd_impl(const b200_convt_fwd_params* p) {
  if (!p || !view_ok(&p->x) || !view_ok(&p->y)) return -1;
  return umma_convt_fwd_ok(p) ? B200_IMPL_UMMA : B200_IMPL_DIRECT;
}
int b200unet_convt_dgrad_impl(const b200_convt_dgrad_params* p) {
  if (!p || !view_ok(&p->dx) || !view_ok(&p->dy)) return -1;
  return umma_convt_dgrad_ok(p) ? B200_IMPL_UMMA : B200_IMPL_DIRECT;
}
int b200unet_convt_wgrad_impl(const b200_convt_wgrad_params* p) {
  if (!p || !view_ok(&p->x) || !view_ok(&p->dy)) return -1;
  return umma_convt_wgrad_ok(p) ? B200_IMPL_UMMA : B200_IMPL_DIRECT;
}

int b200unet_conv_fwd(const b200_conv_fwd_params* p, void* stream) {
  int r = check_conv_fwd(p), impl;
  if (r) return r;
  if ((r = resolve(p, umma_conv_fwd_ok, "conv_fwd", &impl))) return r;
  if (impl == B200_IMPL_UMMA) return umma_conv_fwd(p, as_stream(stream));
  if (p->impl == B200_IMPL_AUTO && smallc_conv_fwd_ok(p)) return smallc_conv_fwd(p, as_stream(stream));
  if (p->impl == B200_IMPL_AUTO) note_fallback("conv_fwd", p->src[0].c);
  B200_REQUIRE(!p->dst.lo, "conv_fwd: the split precision tier runs on the tcgen05 / first-layer kernels only");
  return direct_conv_fwd(p, as_stream(stream));
}

int b200unet_conv_dgrad(const b200_conv_dgrad_params* p, void* stream) {
  int r = check_conv_dgrad(p), impl;
  if (r) return r;
  if ((r = resolve(p, umma_conv_dgrad_ok, "conv_dgrad", &impl))) return r;
  if (impl == B200_IMPL_DIRECT && p->impl == B200_IMPL_AUTO) note_fallback("conv_dgrad", p->dst[0].c);
  return impl == B200_IMPL_UMMA ? umma_conv_dgrad(p, as_stream(stream)) : direct_conv_dgrad(p, as_stream(stream));
}

size_t b200unet_conv_wgrad_workspace_bytes(const b200_conv_wgrad_params* p) {
  if (check_conv_wgrad(p)) return 0;
  const size_t d = direct_wgrad_workspace(view_pixels(p->dz), p->dz.c, cin_of(p), p->taps);
  const size_t u = umma_conv_wgrad_ok(p) ? umma_conv_wgrad_workspace(p) : 0;
  if (p->impl == B200_IMPL_DIRECT) return d;
  if (p->impl == B200_IMPL_UMMA) return u;
  if (umma_conv_wgrad_ok(p)) return u;
  return smallc_conv_wgrad_ok(p) ? smallc_conv_wgrad_workspace(p) : d;
}

int b200unet_conv_wgrad(const b200_conv_wgrad_params* p, void* workspace, size_t workspace_bytes, void* stream) {
  int r = check_conv_wgrad(p), impl;
  if (r) return r;
  if ((r = resolve(p, umma_conv_wgrad_ok, "conv_wgrad", &impl))) return r;
  if (impl == B200_IMPL_UMMA) return umma_conv_wgrad(p, workspace, workspace_bytes, as_stream(stream));
  if (p->impl == B200_IMPL_AUTO && smallc_conv_wgrad_ok(p))
    return smallc_conv_wgrad(p, workspace, workspace_bytes, as_stream(stream));
  return direct_conv_wgrad(p, workspace, workspace_bytes, as_stream(stream));
}

int b200unet_convt_fwd(const b200_convt_fwd_params* p, void* stream) {
  B200_REQUIRE(p && view_ok(&p->x) && view_ok(&p->y), "convt_fwd: bad views");
  B200_REQUIRE(p->y.n == p->x.n && p->y.h == 2 * p->x.h && p->y.w == 2 * p->x.w, "convt_fwd: y must be 2x x");
  B200_REQUIRE((p->x.lo != nullptr) == (p->y.lo != nullptr), "convt_fwd: x and y must be of the same precision tier");
  int impl, r;
  if ((r = resolve(p, umma_convt_fwd_ok, "convt_fwd", &impl))) return r;
  B200_REQUIRE(impl == B200_IMPL_UMMA || !p->y.lo, "convt_fwd: the split precision tier runs on the tcgen05 kernel only");
  return impl == B200_IMPL_UMMA ? umma_convt_fwd(p, as_stream(stream)) : direct_convt_fwd(p, as_stream(stream));
}

int b200unet_convt_dgrad(const b200_convt_dgrad_params* p, void* stream) {
  B200_REQUIRE(p && view_ok(&p->dx) && view_ok(&p->dy), "convt_dgrad: bad views");
  B200_REQUIRE(p->dy.n == p->dx.n && p->dy.h == 2 * p->dx.h && p->dy.w == 2 * p->dx.w, "convt_dgrad: dy must be 2x dx");
  int impl, r;
  if ((r = resolve(p, umma_convt_dgrad_ok, "convt_dgrad", &impl))) return r;
  return impl == B200_IMPL_UMMA ? umma_convt_dgrad(p, as_stream(stream)) : direct_convt_dgrad(p, as_stream(stream));
}

size_t b200unet_convt_wgrad_workspace_bytes(const b200_convt_wgrad_params* p) {
  if (!p || !view_ok(&p->x) || !view_ok(&p->dy)) return 0;
  const size_t d = direct_wgrad_workspace(view_pixels(p->x), p->x.c, p->dy.c, 4);
  const size_t u = umma_convt_wgrad_ok(p) ? umma_convt_wgrad_workspace(p) : 0;
  if (p->impl == B200_IMPL_DIRECT) return d;
  if (p->impl == B200_IMPL_UMMA) return u;
  return umma_convt_wgrad_ok(p) ? u : d;
}

int b200unet_convt_wgrad(const b200_convt_wgrad_params* p, void* workspace, size_t workspace_bytes, void* stream) {
  B200_REQUIRE(p && view_ok(&p->x) && view_ok(&p->dy) && p->dw_f32, "convt_wgrad: bad arguments");
  B200_REQUIRE(p->dy.n == p->x.n && p->dy.h == 2 * p->x.h && p->dy.w == 2 * p->x.w, "convt_wgrad: dy must be 2x x");
  int impl, r;
  if ((r = resolve(p, umma_convt_wgrad_ok, "convt_wgrad", &impl))) return r;
  return impl == B200_IMPL_UMMA ? umma_convt_wgrad(p, workspace, workspace_bytes, as_stream(stream))
                                : direct_convt_wgrad(p, workspace, workspace_bytes, as_stream(stream));
}
}
