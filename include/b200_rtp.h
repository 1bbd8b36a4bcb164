/*
 * b200_rtp.h -- what sits between the encoder and the decoder filters:
 * Annex-B NAL splitting and an RFC 7798 (HEVC over RTP) packetiser / depacketiser.
 *
 * In the reference this is uvgRTP: UvgRTPSender::process hands a whole access unit to
 * media_stream::push_frame (src/media/delivery/uvgrtpsender.cpp:89-118), which splits it at the
 * start codes and sends each NAL as a single-NAL packet or as fragmentation units; on the other side
 * UvgRTPReceiver::receiveHook (src/media/delivery/uvgrtpreceiver.cpp:54-116) gets ONE NAL per frame
 * with a 4-byte start code prepended (RCE_H26X_PREPEND_SC, :87-111) and forwards it to
 * OpenHEVCFilter.  uvgRTP is a network-fetched dependency that is not in the reference tree; this
 * shim restates the wire format from RFC 3550 (12-byte RTP header) and RFC 7798 sections 4.4.1
 * (single NAL unit), 4.4.2 (aggregation packets, receive only) and 4.4.3 (fragmentation units) so
 * that a loopback test can run encoder -> packets -> decoder exactly as a call does, including
 * packet loss.  Pure host code: no GPU involved, no sockets -- packets are byte buffers.
 */
#ifndef B200_RTP_H_
#define B200_RTP_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200_nal_span {
  uint32_t offset;     /* first byte of the NAL header inside the buffer (after the start code) */
  uint32_t length;     /* NAL bytes, trailing zero bytes before the next start code stripped */
} b200_nal_span;

/* Splits an Annex-B buffer at 00 00 01 / 00 00 00 01 start codes.  Returns the number of NAL units
 * found (which may exceed `cap`; only the first `cap` spans are written), -1 on bad arguments. */
int b200_annexb_split(const uint8_t *buf, size_t len, b200_nal_span *out, int cap);

/* Filter::isHEVCIntra / isHEVCInter (src/media/processing/filter.cpp:516-532): the buffer starts
 * with a 4-byte start code followed by an IDR_W_RADL (19) / TRAIL_R (1) NAL. */
int b200_is_hevc_intra(const uint8_t *buf, size_t len);
int b200_is_hevc_inter(const uint8_t *buf, size_t len);

/* ---- sender side ---------------------------------------------------------------------------- */
typedef struct b200_rtp_sender b200_rtp_sender;

/* max_payload: largest RTP payload in bytes (uvgRTP default: 1500 MTU - 20 IP - 8 UDP - 12 RTP = 1460). */
b200_rtp_sender *b200_rtp_sender_new(uint32_t ssrc, int payload_type, int max_payload);
void b200_rtp_sender_free(b200_rtp_sender *s);

/* push_frame: packetises one access unit (Annex-B, as encoder_encode returns it).  Packets are
 * written back to back into `out`; pkt_len[i] receives the size of packet i (12-byte RTP header
 * included).  The marker bit is set on the last packet of the access unit; `rtp_timestamp` is in
 * 90 kHz units.  Returns the number of packets, -1 on bad arguments, -2 when `out` or `pkt_len` is
 * too small (nothing is consumed: sequence numbers do not advance). */
int b200_rtp_push_frame(b200_rtp_sender *s, const uint8_t *au, size_t au_len, uint32_t rtp_timestamp,
                        uint8_t *out, size_t out_cap, uint32_t *pkt_len, int max_packets);
/* Upper bound of the bytes / packets push_frame needs for an access unit of au_len bytes. */
size_t b200_rtp_bound_bytes(const b200_rtp_sender *s, size_t au_len);
int    b200_rtp_bound_packets(const b200_rtp_sender *s, size_t au_len);

/* ---- receiver side -------------------------------------------------------------------------- */
typedef struct b200_rtp_receiver b200_rtp_receiver;

b200_rtp_receiver *b200_rtp_receiver_new(uint32_t expected_ssrc);
void b200_rtp_receiver_free(b200_rtp_receiver *r);

/* Feeds one RTP packet.  Returns the number of complete NAL units now queued (0 while a fragmented
 * NAL is still incomplete), -1 for a malformed packet or a wrong SSRC (uvgrtpreceiver.cpp:68-76;
 * the packet is ignored).  A sequence-number gap discards the fragmented NAL being assembled, as
 * uvgRTP does; b200_rtp_receiver_lost() counts the NAL units dropped that way. */
int b200_rtp_receive(b200_rtp_receiver *r, const uint8_t *pkt, size_t len);
/* Pops the oldest queued NAL as the receiver filter sees it: 4-byte start code + NAL.  Returns its
 * size, 0 when the queue is empty, -2 when `cap` is too small (the NAL stays queued and
 * *rtp_timestamp receives the size needed, so the caller can grow its buffer and call again).
 * Memory is bounded: a fragmented NAL growing beyond 16 MB and NALs arriving while 1024 are queued
 * are dropped and counted by b200_rtp_receiver_lost(). */
int b200_rtp_next_nal(b200_rtp_receiver *r, uint8_t *out, size_t cap, uint32_t *rtp_timestamp, int *marker);
unsigned b200_rtp_receiver_lost(const b200_rtp_receiver *r);

#ifdef __cplusplus
}
#endif
#endif /* B200_RTP_H_ */
