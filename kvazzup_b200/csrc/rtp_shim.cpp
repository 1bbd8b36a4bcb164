// Annex-B NAL splitting and RFC 7798 packetisation (include/b200_rtp.h).  Host-only glue between
// the encoder and decoder boundaries; stands in for what uvgRTP does for the reference
// (src/media/delivery/uvgrtpsender.cpp:89-118, uvgrtpreceiver.cpp:54-116).
#include <string.h>

#include <deque>
#include <new>
#include <vector>

#include "../../include/b200_rtp.h"

namespace {

constexpr int kRtpHeader = 12;
constexpr int kNalHeader = 2;
constexpr int kTypeAp = 48, kTypeFu = 49;

// Walks the start codes of an Annex-B buffer; f(offset, length) per NAL unit.
template <class F>
int for_each_nal(const uint8_t *buf, size_t len, F f)
{
  auto find = [&](size_t from) -> size_t {
    for (size_t k = from; k + 3 <= len; k++)
      if (buf[k] == 0 && buf[k + 1] == 0 && buf[k + 2] == 1) return k;
    return len;
  };
  int n = 0;
  size_t sc = find(0);
  while (sc < len) {
    const size_t start = sc + 3;
    const size_t next = find(start);
    size_t end = next;
    if (next < len)
      while (end > start && buf[end - 1] == 0) end--;      // zero_byte of a 4-byte start code, trailing_zero_8bits
    if (end > start) { f(start, end - start); n++; }
    sc = next;
  }
  return n;
}

void put_rtp_header(uint8_t *p, int pt, int marker, uint16_t seq, uint32_t ts, uint32_t ssrc)
{
  p[0] = 0x80;                                              // V=2, P=0, X=0, CC=0
  p[1] = (uint8_t)((marker ? 0x80 : 0) | (pt & 0x7f));
  p[2] = (uint8_t)(seq >> 8); p[3] = (uint8_t)seq;
  p[4] = (uint8_t)(ts >> 24); p[5] = (uint8_t)(ts >> 16); p[6] = (uint8_t)(ts >> 8); p[7] = (uint8_t)ts;
  p[8] = (uint8_t)(ssrc >> 24); p[9] = (uint8_t)(ssrc >> 16); p[10] = (uint8_t)(ssrc >> 8); p[11] = (uint8_t)ssrc;
}

}  // namespace

struct b200_rtp_sender {
  uint32_t ssrc;
  int pt, max_payload;
  uint16_t seq;
};

// Limits of the receiver: a NAL of a level-6.2 picture stays far below 16 MB; 1024 queued NALs are
// more than 30 access units of a 4K tiled picture.
constexpr size_t kMaxNalBytes = 16u << 20;
constexpr size_t kMaxReadyNals = 1024;

struct b200_rtp_receiver {
  struct Nal { std::vector<uint8_t> bytes; uint32_t ts; int marker; };
  uint32_t ssrc;
  bool have_seq = false, fu_open = false, fu_skipping = false;   // skipping: rest of a NAL already counted as lost
  uint16_t last_seq = 0;
  std::vector<uint8_t> fu;          // NAL being reassembled (start code + header + payload so far)
  uint32_t fu_ts = 0;
  std::deque<Nal> ready;
  unsigned lost = 0;
};

extern "C" {

int b200_annexb_split(const uint8_t *buf, size_t len, b200_nal_span *out, int cap)
{
  if (!buf || (cap > 0 && !out)) return -1;
  int k = 0;
  return for_each_nal(buf, len, [&](size_t off, size_t n) {
    if (k < cap) { out[k].offset = (uint32_t)off; out[k].length = (uint32_t)n; }
    k++;
  });
}

static int nal_type_after_start_code(const uint8_t *buf, size_t len)
{
  if (!buf || len < 6 || buf[0] || buf[1] || buf[2] || buf[3] != 1) return -1;
  return buf[4] >> 1;
}
int b200_is_hevc_intra(const uint8_t *buf, size_t len) { return nal_type_after_start_code(buf, len) == 19; }
int b200_is_hevc_inter(const uint8_t *buf, size_t len) { return nal_type_after_start_code(buf, len) == 1; }

b200_rtp_sender *b200_rtp_sender_new(uint32_t ssrc, int payload_type, int max_payload)
{
  if (max_payload < kNalHeader + 2 || payload_type < 0 || payload_type > 127) return NULL;
  b200_rtp_sender *s = new (std::nothrow) b200_rtp_sender();
  if (!s) return NULL;
  s->ssrc = ssrc; s->pt = payload_type; s->max_payload = max_payload; s->seq = 0;
  return s;
}
void b200_rtp_sender_free(b200_rtp_sender *s) { delete s; }

int b200_rtp_bound_packets(const b200_rtp_sender *s, size_t au_len)
{
  if (!s) return -1;
  // every NAL costs at least 4 bytes of the AU (start code + header), FU payload is max_payload - 3
  return (int)(au_len / 4 + au_len / (size_t)(s->max_payload - 3) + 2);
}
size_t b200_rtp_bound_bytes(const b200_rtp_sender *s, size_t au_len)
{
  if (!s) return 0;
  return au_len + (size_t)b200_rtp_bound_packets(s, au_len) * (kRtpHeader + 3);
}

int b200_rtp_push_frame(b200_rtp_sender *s, const uint8_t *au, size_t au_len, uint32_t ts,
                        uint8_t *out, size_t out_cap, uint32_t *pkt_len, int max_packets)
{
  if (!s || !au || !out || !pkt_len) return -1;
  std::vector<b200_nal_span> nals;
  for_each_nal(au, au_len, [&](size_t off, size_t n) { nals.push_back({(uint32_t)off, (uint32_t)n}); });
  // first pass: size check, so that a failed call leaves the sequence counter untouched
  size_t need_bytes = 0;
  int need_pkts = 0;
  const int fu_chunk = s->max_payload - kNalHeader - 1;
  for (const b200_nal_span &n : nals) {
    if (n.length < (uint32_t)kNalHeader) return -1;
    if ((int)n.length <= s->max_payload) { need_pkts++; need_bytes += kRtpHeader + n.length; continue; }
    const size_t body = n.length - kNalHeader;
    const int k = (int)((body + fu_chunk - 1) / fu_chunk);
    need_pkts += k;
    need_bytes += (size_t)k * (kRtpHeader + kNalHeader + 1) + body;
  }
  if (need_pkts > max_packets || need_bytes > out_cap) return -2;
  uint8_t *w = out;
  int np = 0;
  for (size_t i = 0; i < nals.size(); i++) {
    const uint8_t *nal = au + nals[i].offset;
    const uint32_t n = nals[i].length;
    const bool last_nal = i + 1 == nals.size();
    if ((int)n <= s->max_payload) {                          // single NAL unit packet (RFC 7798 4.4.1)
      put_rtp_header(w, s->pt, last_nal, s->seq++, ts, s->ssrc);
      memcpy(w + kRtpHeader, nal, n);
      pkt_len[np++] = kRtpHeader + n;
      w += kRtpHeader + n;
      continue;
    }
    // fragmentation units (4.4.3): PayloadHdr = NAL header with type 49, FU header = S | E | FuType
    const int type = (nal[0] >> 1) & 63;
    const uint8_t *body = nal + kNalHeader;
    size_t left = n - kNalHeader;
    bool first = true;
    while (left) {
      const size_t c = left < (size_t)fu_chunk ? left : (size_t)fu_chunk;
      const bool end = c == left;
      put_rtp_header(w, s->pt, last_nal && end, s->seq++, ts, s->ssrc);
      uint8_t *p = w + kRtpHeader;
      p[0] = (uint8_t)((nal[0] & 0x81) | (kTypeFu << 1));
      p[1] = nal[1];
      p[2] = (uint8_t)((first ? 0x80 : 0) | (end ? 0x40 : 0) | type);
      memcpy(p + 3, body, c);
      pkt_len[np++] = (uint32_t)(kRtpHeader + 3 + c);
      w += kRtpHeader + 3 + c;
      body += c; left -= c; first = false;
    }
  }
  return np;
}

b200_rtp_receiver *b200_rtp_receiver_new(uint32_t expected_ssrc)
{
  b200_rtp_receiver *r = new (std::nothrow) b200_rtp_receiver();
  if (r) r->ssrc = expected_ssrc;
  return r;
}
void b200_rtp_receiver_free(b200_rtp_receiver *r) { delete r; }
unsigned b200_rtp_receiver_lost(const b200_rtp_receiver *r) { return r ? r->lost : 0; }

int b200_rtp_receive(b200_rtp_receiver *r, const uint8_t *pkt, size_t len)
{
  if (!r || !pkt || len < (size_t)kRtpHeader + kNalHeader || (pkt[0] >> 6) != 2) return -1;
  const int cc = pkt[0] & 15;
  size_t hdr = kRtpHeader + 4 * (size_t)cc;
  if (pkt[0] & 0x10) {                                       // header extension
    if (len < hdr + 4) return -1;
    hdr += 4 + 4 * (((size_t)pkt[hdr + 2] << 8) | pkt[hdr + 3]);
  }
  if (len < hdr + kNalHeader) return -1;                     // CSRC list / extension longer than the packet (also keeps len - hdr below from wrapping)
  size_t end = len;
  if (pkt[0] & 0x20) {                                       // padding
    if (pkt[len - 1] == 0 || pkt[len - 1] > len - hdr) return -1;
    end -= pkt[len - 1];
  }
  if (end < hdr + kNalHeader) return -1;
  const uint32_t ssrc = ((uint32_t)pkt[8] << 24) | ((uint32_t)pkt[9] << 16) | ((uint32_t)pkt[10] << 8) | pkt[11];
  if (ssrc != r->ssrc) return -1;
  const int marker = pkt[1] >> 7;
  const uint16_t seq = (uint16_t)((pkt[2] << 8) | pkt[3]);
  const uint32_t ts = ((uint32_t)pkt[4] << 24) | ((uint32_t)pkt[5] << 16) | ((uint32_t)pkt[6] << 8) | pkt[7];
  const uint8_t *p = pkt + hdr;
  const size_t n = end - hdr;
  const int type = (p[0] >> 1) & 63;
  if (type >= 50 || (type == kTypeFu && n < 4)) return -1;   // PACI / reserved / truncated FU: sequence state untouched
  const bool gap = r->have_seq && seq != (uint16_t)(r->last_seq + 1);
  r->have_seq = true; r->last_seq = seq;
  // Bounded memory: a sender that never sets the E bit, or a consumer that never pops, must not
  // grow this receiver without limit.  Oversized NALs and NALs arriving at a full queue are dropped
  // and counted as lost.
  if (r->ready.size() >= kMaxReadyNals) { r->ready.pop_front(); r->lost++; }
  if (gap && r->fu_open) { r->fu_open = false; r->fu_skipping = true; r->fu.clear(); r->lost++; }
  static const uint8_t sc[4] = {0, 0, 0, 1};
  if (type == kTypeFu) {
    const bool s_bit = p[2] & 0x80, e_bit = p[2] & 0x40;
    if (s_bit) {
      if (r->fu_open) r->lost++;                             // previous NAL never saw its end fragment
      r->fu.assign(sc, sc + 4);
      r->fu.push_back((uint8_t)((p[0] & 0x81) | ((p[2] & 63) << 1)));
      r->fu.push_back(p[1]);
      r->fu_open = true; r->fu_skipping = false; r->fu_ts = ts;
    } else if (!r->fu_open) {
      if (!r->fu_skipping) { r->lost++; r->fu_skipping = true; }   // fragment of a NAL whose start was lost
      if (e_bit) r->fu_skipping = false;
      return (int)r->ready.size();
    }
    if (r->fu.size() + (n - 3) > kMaxNalBytes) {             // runaway fragment train: drop the NAL, skip to its end
      r->fu.clear(); r->fu.shrink_to_fit(); r->fu_open = false; r->fu_skipping = !e_bit; r->lost++;
      return (int)r->ready.size();
    }
    r->fu.insert(r->fu.end(), p + 3, p + n);
    if (e_bit) {
      r->ready.push_back({std::move(r->fu), r->fu_ts, marker});
      r->fu.clear(); r->fu_open = false;
    }
    return (int)r->ready.size();
  }
  if (r->fu_open) { r->fu_open = false; r->fu.clear(); r->lost++; }
  r->fu_skipping = false;
  if (type == kTypeAp) {                                     // aggregation packet (4.4.2): 16-bit size + NAL, repeated
    size_t off = kNalHeader;
    std::vector<b200_rtp_receiver::Nal> got;
    while (off + 2 <= n) {
      const size_t sz = ((size_t)p[off] << 8) | p[off + 1];
      off += 2;
      if (sz < (size_t)kNalHeader || off + sz > n) return -1;
      b200_rtp_receiver::Nal nal{std::vector<uint8_t>(sc, sc + 4), ts, 0};
      nal.bytes.insert(nal.bytes.end(), p + off, p + off + sz);
      got.push_back(std::move(nal));
      off += sz;
    }
    if (got.empty() || off != n) return -1;
    got.back().marker = marker;
    for (auto &g : got) r->ready.push_back(std::move(g));
    return (int)r->ready.size();
  }
  b200_rtp_receiver::Nal nal{std::vector<uint8_t>(sc, sc + 4), ts, marker};
  nal.bytes.insert(nal.bytes.end(), p, p + n);
  r->ready.push_back(std::move(nal));
  return (int)r->ready.size();
}

int b200_rtp_next_nal(b200_rtp_receiver *r, uint8_t *out, size_t cap, uint32_t *ts, int *marker)
{
  if (!r || !out) return -1;
  if (r->ready.empty()) return 0;
  const b200_rtp_receiver::Nal &n = r->ready.front();
  if (n.bytes.size() > cap) {                                // caller's buffer too small: report the size needed, keep the NAL
    if (ts) *ts = (uint32_t)n.bytes.size();
    return -2;
  }
  memcpy(out, n.bytes.data(), n.bytes.size());
  if (ts) *ts = n.ts;
  if (marker) *marker = n.marker;
  const int len = (int)n.bytes.size();
  r->ready.pop_front();
  return len;
}

}  // extern "C"
