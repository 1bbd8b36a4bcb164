// ASan / UBSan harness over the RTP shim (rtp_shim.cpp, host only): packetises random access units, then feeds
// the receiver the packets mutated, truncated, reordered, duplicated and mixed with random datagrams, and pops
// NAL units into buffers of random sizes.  Deterministic (xorshift seeded from argv[1]).
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "b200_rtp.h"

static uint64_t s_state = 88172645463325252ull;
static uint32_t rnd() { s_state ^= s_state << 13; s_state ^= s_state >> 7; s_state ^= s_state << 17; return (uint32_t)(s_state >> 16); }

int main(int argc, char **argv)
{
  if (argc > 1) s_state ^= strtoull(argv[1], nullptr, 10) * 0x9e3779b97f4a7c15ull;
  const int rounds = argc > 2 ? atoi(argv[2]) : 2000;
  unsigned long pkts = 0, nals = 0, refused = 0;
  for (int round = 0; round < rounds; round++) {
    const int max_payload = (rnd() % 8 == 0) ? (int)(rnd() % 64) : 100 + (int)(rnd() % 1400);
    b200_rtp_sender *tx = b200_rtp_sender_new(0x1234, 96, max_payload);
    b200_rtp_receiver *rx = b200_rtp_receiver_new(rnd() % 4 ? 0x1234 : 0);
    if (!tx || !rx) { if (tx) b200_rtp_sender_free(tx); if (rx) b200_rtp_receiver_free(rx); refused++; continue; }
    for (int frame = 0; frame < 6; frame++) {
      // an access unit: a few NAL units of random sizes with 3- or 4-byte start codes, bytes that look like start codes inside
      std::vector<uint8_t> au;
      const int n_nal = 1 + rnd() % 5;
      for (int k = 0; k < n_nal; k++) {
        if (rnd() & 1) au.push_back(0);
        au.push_back(0); au.push_back(0); au.push_back(1);
        const int type = rnd() % 4 == 0 ? 32 + rnd() % 3 : (rnd() & 1 ? 19 : 1);
        au.push_back((uint8_t)(type << 1)); au.push_back(1);
        const size_t len = rnd() % 8 == 0 ? rnd() % 40000 : rnd() % 3000;
        for (size_t i = 0; i < len; i++) au.push_back((uint8_t)(rnd() % 7 ? rnd() : 0));
      }
      if (rnd() % 16 == 0) au.resize(rnd() % (au.size() + 1));
      const size_t cap_b = b200_rtp_bound_bytes(tx, au.size());
      const int cap_p = b200_rtp_bound_packets(tx, au.size());
      std::vector<uint8_t> out(rnd() % 20 == 0 ? cap_b / 2 : cap_b);
      std::vector<uint32_t> pl((size_t)(rnd() % 20 == 0 ? cap_p / 2 : cap_p) + 1);
      const int n = b200_rtp_push_frame(tx, au.data(), au.size(), 3000u * frame, out.data(), out.size(), pl.data(), (int)pl.size() - 1);
      if (n < 0) { refused++; continue; }
      size_t off = 0;
      std::vector<std::vector<uint8_t>> packets;
      for (int i = 0; i < n; i++) { packets.emplace_back(out.begin() + off, out.begin() + off + pl[i]); off += pl[i]; }
      for (size_t i = 0; i < packets.size(); i++) {
        std::vector<uint8_t> p = packets[i];
        const uint32_t m = rnd() % 16;
        if (m == 0) continue;                                                 // lost
        if (m == 1 && !p.empty()) p.resize(rnd() % (p.size() + 1));           // truncated
        if (m == 2) for (int k = 0; k < 4 && !p.empty(); k++) p[rnd() % p.size()] ^= (uint8_t)(1u << (rnd() % 8));
        if (m == 3) { p.resize(rnd() % 64); for (auto &b : p) b = (uint8_t)rnd(); }   // a random datagram
        if (m == 4 && i + 1 < packets.size()) std::swap(p, packets[i + 1]);    // reordered
        if (m == 5 && p.size() > 14) { p[12] = (uint8_t)(48 << 1); }           // pretend it is an aggregation packet
        if (m == 6 && p.size() > 14) { p[12] = (uint8_t)(49 << 1); p[14] = (uint8_t)rnd(); }   // ... a fragmentation unit with a random FU header
        b200_rtp_receive(rx, p.data(), p.size());
        if (m == 7) b200_rtp_receive(rx, p.data(), p.size());                  // duplicated
        pkts++;
      }
      for (;;) {
        std::vector<uint8_t> nal(rnd() % 4 == 0 ? rnd() % 200 : 70000);
        uint32_t ts = 0; int marker = 0;
        int got = b200_rtp_next_nal(rx, nal.data(), nal.size(), &ts, &marker);
        if (got == -2) { nal.resize(ts); got = b200_rtp_next_nal(rx, nal.data(), nal.size(), &ts, &marker); }
        if (got <= 0) break;
        nals++;
        b200_is_hevc_intra(nal.data(), (size_t)got); b200_is_hevc_inter(nal.data(), (size_t)got);
        b200_nal_span spans[4];
        b200_annexb_split(nal.data(), (size_t)got, spans, 4);
      }
    }
    b200_rtp_receiver_lost(rx);
    b200_rtp_sender_free(tx);
    b200_rtp_receiver_free(rx);
  }
  printf("rounds %d packets %lu nals %lu refused %lu\n", rounds, pkts, nals, refused);
  return 0;
}
