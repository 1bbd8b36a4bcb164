// ASan / UBSan harness over the host pass of the MJPG path (mjpg.cu: marker parsing, Huffman tables, the entropy-coded
// scan) through b200_mjpg_probe, which runs it without touching the GPU.  Input: records {u32 length, bytes}.
#include <stdint.h>
#include <stdio.h>
#include <vector>
#include "b200media.h"

int main(int argc, char **argv)
{
  FILE *f = fopen(argv[1], "rb");
  if (!f) return 2;
  unsigned long cases = 0, ok = 0;
  for (;;) {
    uint32_t n;
    if (fread(&n, 4, 1, f) != 1) break;
    std::vector<uint8_t> buf(n);
    if (n && fread(buf.data(), 1, n, f) != n) break;
    int w = 0, h = 0, ss = 0;
    cases++;
    if (b200_mjpg_probe(buf.data(), buf.size(), &w, &h, &ss) == 0) ok++;
  }
  printf("cases %lu accepted %lu\n", cases, ok);
  return 0;
}
