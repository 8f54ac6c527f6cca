// tools/pack_timing.cc -- host-only timing of the pack phases of tba_upload (theiasfm_b200/csrc/tba_pack.h) on a synthetic
// problem read from a raw dump (written by tools/pack_timing.py).  Not part of the product; used to tune the host pack.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>
#ifndef PACK_HEADER
#define PACK_HEADER "../theiasfm_b200/csrc/tba_pack.h"
#endif
#include PACK_HEADER
using namespace tba;
static uint64_t fnv(const void* d, size_t n, uint64_t h) { const unsigned char* b = (const unsigned char*)d; for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; } return h; }
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
template <class T> static std::vector<T> rd(FILE* f, size_t n) { std::vector<T> v(n); if (fread(v.data(), sizeof(T), n, f) != n) { perror("read"); exit(1); } return v; }
int main(int argc, char** argv) {
  if (argc < 3) { fprintf(stderr, "usage: pack_timing dump.bin threads [reps]\n"); return 2; }
  FILE* f = fopen(argv[1], "rb");
  const int T = atoi(argv[2]), reps = argc > 3 ? atoi(argv[3]) : 3;
  int64_t hdr[4];
  if (fread(hdr, 8, 4, f) != 4) return 1;
  const int nc = (int)hdr[0], ng = (int)hdr[1], np = (int)hdr[2]; const int64_t no = hdr[3];
  auto ext = rd<double>(f, (size_t)nc * 6); auto extc = rd<uint8_t>(f, nc); auto cg = rd<int32_t>(f, nc); auto gm = rd<int32_t>(f, ng);
  auto intr = rd<double>(f, (size_t)ng * 10); auto mk = rd<uint32_t>(f, ng); auto pt = rd<double>(f, (size_t)np * 4); auto ptc = rd<uint8_t>(f, np);
  auto oc = rd<int32_t>(f, no); auto op = rd<int32_t>(f, no); auto xy = rd<double>(f, (size_t)no * 2);
  tba_problem p{}; p.n_cam = nc; p.ext = ext.data(); p.ext_const = extc.data(); p.cam_group = cg.data(); p.n_group = ng; p.group_model = gm.data();
  p.intr = intr.data(); p.group_const_mask = mk.data(); p.n_pt = np; p.pt = pt.data(); p.pt_const = ptc.data(); p.n_obs = no; p.obs_cam = oc.data();
  p.obs_pt = op.data(); p.obs_xy = xy.data();
  std::vector<double> dxy, dpt; std::vector<int> scam, spt; std::vector<int16_t> srun; std::vector<uint8_t> sfl, pc; std::vector<int64_t> so;
  HostPack H;
  for (int r = 0; r < reps; ++r) {
    const double t0 = now();
    pack_count_and_sort(&p, T, &H);
    const double t1 = now();
    pack_points(&p, &H);
    std::vector<double> cnt_c(nc, 0.0), cnt_g(ng, 0.0);
    for (int i = 0; i < nc; ++i) { cnt_c[i] = H.cnt_cam[i]; cnt_g[cg[i]] += H.cnt_cam[i]; }
    pack_masks_and_tiles(&p, cnt_c, cnt_g, &H);
    const double t2 = now();
    dxy.resize((size_t)H.n_slots * 2); dpt.resize(H.pk2caller.size() * 4); scam.resize(H.n_slots); spt.resize(H.n_slots); srun.resize(H.n_slots);
    sfl.resize(H.n_slots); pc.resize(H.pk2caller.size());
    const bool want_orig = r == reps - 1 || getenv("PACK_OLD") != nullptr;
    if (want_orig) so.assign(H.n_slots, -1);
    const double t3 = now();
    PackDest d; d.xy = dxy.data(); d.pt = dpt.data(); d.slot_cam = scam.data(); d.slot_pt = spt.data(); d.slot_run = srun.data(); d.slot_flags = sfl.data();
    d.pt_const = pc.data(); d.slot_orig = want_orig ? so.data() : nullptr;
    pack_fill(&p, H, T, d);
    const double t4 = now();
    if (r == reps - 1) {
      uint64_t h = 1469598103934665603ull;
      h = fnv(dxy.data(), dxy.size() * 8, h); h = fnv(dpt.data(), dpt.size() * 8, h); h = fnv(scam.data(), scam.size() * 4, h); h = fnv(spt.data(), spt.size() * 4, h);
      h = fnv(srun.data(), srun.size() * 2, h); h = fnv(sfl.data(), sfl.size(), h); h = fnv(pc.data(), pc.size(), h); h = fnv(so.data(), so.size() * 8, h);
      h = fnv(H.pk2caller.data(), H.pk2caller.size() * 4, h); h = fnv(H.tile_pt_begin.data(), H.tile_pt_begin.size() * 4, h);
      h = fnv(H.tile_nruns.data(), H.tile_nruns.size() * 4, h); h = fnv(H.tile_flags.data(), H.tile_flags.size(), h); h = fnv(H.mask.data(), H.mask.size() * 8, h);
      printf("output hash %016llx\n", (unsigned long long)h);
    }
    printf("rep %d: count+sort %.1f ms, points+tiles %.1f ms, alloc/slot_orig %.1f ms, fill %.1f ms, total %.1f ms (slots %lld, tiles %d)\n", r, 1e3 * (t1 - t0),
           1e3 * (t2 - t1), 1e3 * (t3 - t2), 1e3 * (t4 - t3), 1e3 * (t4 - t0), (long long)H.n_slots, H.n_tiles);
  }
  return 0;
}
