// ASan / UBSan harness over the host-side HEVC header parsers (hevc_headers.cpp): reads records
// {u32 length, bytes} of Annex-B buffers and walks them like b200_dec_probe does.
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>
#include "hevc_headers.h"
using namespace b200;
static std::vector<uint8_t> unescape(const uint8_t *p, size_t n)
{
  std::vector<uint8_t> o; o.reserve(n); int z = 0;
  for (size_t i = 0; i < n; i++) { if (z >= 2 && p[i] == 3) { z = 0; continue; } o.push_back(p[i]); z = p[i] == 0 ? z + 1 : 0; }
  return o;
}
int main(int argc, char **argv)
{
  FILE *f = fopen(argv[1], "rb");
  std::vector<Sps> sps(16); std::vector<Pps> pps(64);
  unsigned long ok = 0, bad = 0, cases = 0;
  for (;;) {
    uint32_t n;
    if (fread(&n, 4, 1, f) != 1) break;
    std::vector<uint8_t> buf(n);
    if (fread(buf.data(), 1, n, f) != n) break;
    cases++;
    for (auto &s : sps) s = Sps();
    for (auto &p : pps) p = Pps();
    auto next_sc = [&](size_t from) { for (size_t k = from; k + 3 <= n; k++) if (buf[k] == 0 && buf[k + 1] == 0 && buf[k + 2] == 1) return k; return (size_t)n; };
    for (size_t pos = next_sc(0); pos < n;) {
      const size_t start = pos + 3, next = next_sc(start);
      size_t end = next;
      while (end > start && next < n && buf[end - 1] == 0) end--;
      if (end - start >= 2) {
        const int type = (buf[start] >> 1) & 63;
        std::vector<uint8_t> r = unescape(buf.data() + start + 2, end - start - 2);
        std::string err;
        if (type == 33) { Sps t; if (parse_sps_rbsp(r.data(), r.size(), t, err) && t.id <= 15) { sps[t.id] = t; ok++; } else bad++; }
        if (type == 34) { Pps t; if (parse_pps_rbsp(r.data(), r.size(), t, err)) { pps[t.id] = t; ok++; } else bad++; }
        if (type <= 9 || (type >= 16 && type <= 21)) {
          BitReader pb(r.data(), r.size());
          pb.u(1); if (type >= 16 && type <= 23) pb.u(1);
          const uint32_t pid = pb.ue();
          if (!pb.bad && pid <= 63 && pps[pid].valid && pps[pid].sps_id <= 15 && sps[pps[pid].sps_id].valid) {
            SliceHeader sh;
            if (parse_slice_header_rbsp(r.data(), r.size(), type, sps[pps[pid].sps_id], pps[pid], sh, err)) ok++; else bad++;
          }
        }
      }
      pos = next;
    }
  }
  printf("cases %lu parsed %lu refused %lu\n", cases, ok, bad);
  return 0;
}
