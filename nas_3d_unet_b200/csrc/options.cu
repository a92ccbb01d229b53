// One process-wide table of kernel-selection options, read ONCE (environment NAS3D_<NAME> at
// library load, for A/B runs of the bench) and changed afterwards only through the C-ABI
// (nas3d_set_option) - no getenv on any launch path.  Every option selects between kernels that
// compute the same result (parity tests flip them and compare); none is a fallback off the GPU.
#include "common.cuh"

#include <string.h>

namespace nas3d {

Options g_opt;

struct OptEntry { const char* name; int Options::*field; int lo, hi; };
static const OptEntry kOptions[] = {
    {"tiled", &Options::tiled, 0, 1},
    {"tma", &Options::tma, 0, 1},
    {"tma_merged", &Options::tma_merged, 0, 1},
    {"affine_ring", &Options::affine_ring, 0, 1},
    {"apply_ring", &Options::apply_ring, 0, 1},
    {"reduce_ring", &Options::reduce_ring, 0, 1},
    {"pw_fwd_ring", &Options::pw_fwd_ring, 0, 1},
    {"reduce_waves", &Options::reduce_waves, 0, 1},
    {"ring_min_log2", &Options::ring_min_log2, 10, 40},
    {"pw_vpt_sfb", &Options::pw_vpt_sfb, 1, 4},
    {"pw_vpt_bfs", &Options::pw_vpt_bfs, 1, 4},
    {"pw_vpt_mom", &Options::pw_vpt_mom, 1, 4},
    {"umma_split_k", &Options::umma_split_k, 0, 1},
    {"umma_wgrad", &Options::umma_wgrad, 0, 1},
    {"umma_wgrad_min_c", &Options::umma_wgrad_min_c, 16, 128},
    {"umma_ws", &Options::umma_ws, 0, 1},
    {"s1_wgrad_tma", &Options::s1_wgrad_tma, 0, 1},
    {"s2_wgrad_tma", &Options::s2_wgrad_tma, 0, 1},
};

static const OptEntry* find_option(const char* name) {
  if (!name) return nullptr;
  for (const OptEntry& e : kOptions)
    if (strcmp(e.name, name) == 0) return &e;
  return nullptr;
}

static int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// environment overrides, once: NAS3D_TILED=0, NAS3D_APPLY_RING=0, ... (upper-cased option names)
static const bool g_env_applied = [] {
  for (const OptEntry& e : kOptions) {
    char key[64] = "NAS3D_";
    size_t n = strlen(key);
    for (const char* p = e.name; *p && n + 1 < sizeof(key); ++p)
      key[n++] = (*p >= 'a' && *p <= 'z') ? (char)(*p - 'a' + 'A') : *p;
    key[n] = 0;
    const char* v = getenv(key);
    if (v && *v) g_opt.*(e.field) = clampi(atoi(v), e.lo, e.hi);
  }
  return true;
}();

}  // namespace nas3d

using namespace nas3d;

extern "C" {

int nas3d_set_option(const char* name, int value) {
  const OptEntry* e = find_option(name);
  if (!e) return fail(NAS3D_ERR_ARG, "nas3d_set_option: unknown option '%s'", name ? name : "(null)");
  if (value < e->lo || value > e->hi)
    return fail(NAS3D_ERR_ARG, "nas3d_set_option: %s = %d outside [%d, %d]", name, value, e->lo, e->hi);
  g_opt.*(e->field) = value;
  return NAS3D_OK;
}

int nas3d_get_option(const char* name) {
  const OptEntry* e = find_option(name);
  if (!e) return fail(NAS3D_ERR_ARG, "nas3d_get_option: unknown option '%s'", name ? name : "(null)");
  return g_opt.*(e->field);
}

}  // extern "C"
