// BASELINE config 4 in the reference's own shape: N participant streams on one GPU, one host thread
// per stream, each thread doing exactly what KvazaarFilter::feedInput / parseEncodedFrame
// (kvazaarfilter.cpp:374-484), the uvgRTP hand-over (one NAL per buffer) and
// OpenHEVCFilter::process / sendDecodedOutput (openhevcfilter.cpp:103-239) do -- in C++ over the
// C ABI of libb200media.so, no Python in the loop.
//
//   conference_bench <frames.yuv> <w> <h> <frames_in_file> <streams> <frames_per_stream> <encode_only> <decoder_frame_threads> [pace_fps] [n_devices]
//
// n_devices > 1: participant stream s lives on GPU s mod n_devices (encoder via "b200-device",
// decoder via the thread's current device) -- the streams share no data, so nothing crosses GPUs
// (SURVEY.md 8e).  The FNV-1a hash of every stream's access units is printed: it must not depend on
// n_devices (determinism check).
//
// pace_fps > 0: every stream delivers its pictures at that rate (a live call) with nothing in flight
// (owf 0) and the time from handing a picture to the encoder until its decoded copy is back in host
// memory is recorded per picture: the glass-to-glass share of the media path under N-stream load.
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <thread>
#include <vector>

#include "b200_kvazaar.h"
#include "b200_openhevc.h"
#include "b200_rtp.h"
#include "b200media.h"

static int W, H, NFILE, STREAMS, FRAMES, ENC_ONLY, DEC_THREADS, NDEV = 1;
static std::vector<uint64_t> g_hash;               // per stream: FNV-1a over all access units of the timed pictures
static double PACE_FPS = 0;
static std::vector<std::vector<float>> g_lat;      // per stream: latency of every picture, ms
static std::vector<uint8_t> g_yuv;
static pthread_barrier_t g_bar;
static std::vector<int> g_decoded, g_ok;

struct Stream {
  const kvz_api *api = nullptr;
  kvz_config *cfg = nullptr;
  kvz_encoder *enc = nullptr;
  std::vector<kvz_picture *> pics;
  int next = 0;
  OpenHevc_Handle dec = nullptr;
  std::vector<uint8_t> au, nal, out;
  int decoded = 0;
  uint64_t hash = 1469598103934665603ull;
};

static bool decode_au(Stream &s)
{
  b200_nal_span spans[16];
  int n = b200_annexb_split(s.au.data(), s.au.size(), spans, 16);
  for (int i = 0; i < n && i < 16; i++) {
    s.nal.resize(4 + spans[i].length);
    s.nal[0] = s.nal[1] = s.nal[2] = 0; s.nal[3] = 1;                       // uvgrtpreceiver.cpp:87-111
    memcpy(s.nal.data() + 4, s.au.data() + spans[i].offset, spans[i].length);
    int got = libOpenHevcDecode(s.dec, s.nal.data(), (int)s.nal.size(), 0);
    if (got < 0) { fprintf(stderr, "decode: %s\n", b200_last_error()); return false; }
    if (got > 0) {
      OpenHevc_Frame fr;
      if (libOpenHevcGetOutput(s.dec, got, &fr) > 0) {                      // sendDecodedOutput: repack to packed I420
        libOpenHevcGetPictureInfo(s.dec, &fr.frameInfo);
        const int w = fr.frameInfo.nWidth, h = fr.frameInfo.nHeight;
        s.out.resize((size_t)w * h * 3 / 2);
        for (int r = 0; r < h; r++) memcpy(&s.out[(size_t)r * w], (uint8_t *)fr.pvY + (size_t)r * fr.frameInfo.nYPitch, w);
        for (int r = 0; r < h / 2; r++) {
          memcpy(&s.out[(size_t)w * h + (size_t)r * (w / 2)], (uint8_t *)fr.pvU + (size_t)r * fr.frameInfo.nUPitch, w / 2);
          memcpy(&s.out[(size_t)w * h * 5 / 4 + (size_t)r * (w / 2)], (uint8_t *)fr.pvV + (size_t)r * fr.frameInfo.nVPitch, w / 2);
        }
        s.decoded++;
      }
    }
  }
  return true;
}

static bool feed(Stream &s, const uint8_t *frame)
{
  kvz_data_chunk *chunks = nullptr;
  uint32_t len = 0;
  kvz_picture *recon = nullptr;
  kvz_frame_info info;
  kvz_picture *pic = nullptr;
  if (frame) {
    pic = s.pics[s.next];
    s.next = (s.next + 1) % (int)s.pics.size();
    memcpy(pic->y, frame, (size_t)W * H);                                    // kvazaarfilter.cpp:410-418
    memcpy(pic->u, frame + (size_t)W * H, (size_t)W * H / 4);
    memcpy(pic->v, frame + (size_t)W * H * 5 / 4, (size_t)W * H / 4);
  }
  if (s.api->encoder_encode(s.enc, pic, &chunks, &len, &recon, nullptr, &info) != 1) { fprintf(stderr, "encode: %s\n", b200_last_error()); return false; }
  if (!chunks) return true;
  s.au.clear();
  for (kvz_data_chunk *c = chunks; c; c = c->next) s.au.insert(s.au.end(), c->data, c->data + c->len);   // :465-474
  s.api->chunk_free(chunks);
  s.api->picture_free(recon);
  for (uint8_t b : s.au) { s.hash ^= b; s.hash *= 1099511628211ull; }
  if (!ENC_ONLY) return decode_au(s);
  return true;
}

static void *worker(void *arg)
{
  const int sid = (int)(intptr_t)arg;
  Stream s;
  s.api = kvz_api_get(8);
  s.cfg = s.api->config_alloc();
  s.api->config_init(s.cfg);
  s.api->config_parse(s.cfg, "preset", "veryfast");
  s.cfg->width = W; s.cfg->height = H; s.cfg->framerate_num = 30; s.cfg->framerate_denom = 1;
  s.api->config_parse(s.cfg, "qp", "32");
  s.api->config_parse(s.cfg, "period", "64");
  s.api->config_parse(s.cfg, "owf", PACE_FPS > 0 ? "0" : "3");
  const int dev = sid % NDEV;
  char devs[16];
  snprintf(devs, sizeof(devs), "%d", dev);
  s.api->config_parse(s.cfg, "b200-device", devs);
  b200_set_device(dev);                                                     // the decoder lives on the thread's current device
  s.enc = s.api->encoder_open(s.cfg);
  bool ok = s.enc != nullptr;
  for (int i = 0; ok && i < s.cfg->owf + 1; i++) s.pics.push_back(s.api->picture_alloc(W, H));
  if (ok && !ENC_ONLY) {
    s.dec = libOpenHevcInit(DEC_THREADS, DEC_THREADS > 1 ? 1 : 2);
    ok = libOpenHevcStartDecoder(s.dec) != -1;
  }
  const size_t fb = (size_t)W * H * 3 / 2;
  for (int t = 0; ok && t < 6; t++) ok = feed(s, &g_yuv[(size_t)((t + sid) % NFILE) * fb]);
  pthread_barrier_wait(&g_bar);
  s.decoded = 0;
  s.hash = 1469598103934665603ull;
  const auto t_start = std::chrono::steady_clock::now();
  for (int t = 0; ok && t < FRAMES; t++) {
    if (PACE_FPS > 0) {
      // streams are phase-shifted across the frame interval like independent cameras
      const double due = (t + (double)sid / STREAMS) / PACE_FPS;
      std::this_thread::sleep_until(t_start + std::chrono::duration_cast<std::chrono::steady_clock::duration>(std::chrono::duration<double>(due)));
    }
    const auto t0 = std::chrono::steady_clock::now();
    const int before = s.decoded;
    ok = feed(s, &g_yuv[(size_t)((t + 6 + sid) % NFILE) * fb]);
    if (PACE_FPS > 0 && (ENC_ONLY || s.decoded > before))
      g_lat[sid].push_back(std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count());
  }
  for (int i = 0; ok && i < s.cfg->owf; i++) ok = feed(s, nullptr);          // drain the encoder pipeline
  if (ok && s.dec) {
    int got;
    while ((got = libOpenHevcDecode(s.dec, nullptr, 0, 0)) > 0) s.decoded++; // held-back pictures of frame threading
  }
  pthread_barrier_wait(&g_bar);
  g_decoded[sid] = s.decoded;
  g_ok[sid] = ok;
  g_hash[sid] = s.hash;
  if (s.dec) libOpenHevcClose(s.dec);
  for (kvz_picture *p : s.pics) s.api->picture_free(p);
  if (s.enc) s.api->encoder_close(s.enc);
  s.api->config_destroy(s.cfg);
  return nullptr;
}

int main(int argc, char **argv)
{
  if (argc < 9) { fprintf(stderr, "usage: %s frames.yuv w h frames_in_file streams frames_per_stream encode_only decoder_frame_threads\n", argv[0]); return 2; }
  W = atoi(argv[2]); H = atoi(argv[3]); NFILE = atoi(argv[4]); STREAMS = atoi(argv[5]); FRAMES = atoi(argv[6]);
  ENC_ONLY = atoi(argv[7]); DEC_THREADS = atoi(argv[8]);
  if (argc > 9) PACE_FPS = atof(argv[9]);
  if (argc > 10) NDEV = std::max(1, atoi(argv[10]));
  const size_t fb = (size_t)W * H * 3 / 2;
  g_yuv.resize(fb * NFILE);
  FILE *f = fopen(argv[1], "rb");
  if (!f || fread(g_yuv.data(), 1, g_yuv.size(), f) != g_yuv.size()) { fprintf(stderr, "cannot read %s\n", argv[1]); return 2; }
  fclose(f);
  g_decoded.assign(STREAMS, 0); g_ok.assign(STREAMS, 0);
  g_lat.assign(STREAMS, {});
  g_hash.assign(STREAMS, 0);
  pthread_barrier_init(&g_bar, nullptr, STREAMS + 1);
  std::vector<pthread_t> th(STREAMS);
  for (int i = 0; i < STREAMS; i++) pthread_create(&th[i], nullptr, worker, (void *)(intptr_t)i);
  pthread_barrier_wait(&g_bar);
  auto t0 = std::chrono::steady_clock::now();
  pthread_barrier_wait(&g_bar);
  double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  for (auto &t : th) pthread_join(t, nullptr);
  bool all = true;
  for (int i = 0; i < STREAMS; i++) all = all && g_ok[i] && (ENC_ONLY || g_decoded[i] >= FRAMES);
  double fps = (double)STREAMS * FRAMES / dt;
  uint64_t all_hash = 1469598103934665603ull;                               // over the per-stream hashes, in stream order
  for (int i = 0; i < STREAMS; i++) for (int b = 0; b < 8; b++) { all_hash ^= (g_hash[i] >> (8 * b)) & 0xff; all_hash *= 1099511628211ull; }
  if (PACE_FPS > 0) {
    std::vector<float> lat;
    for (auto &v : g_lat) lat.insert(lat.end(), v.begin(), v.end());
    std::sort(lat.begin(), lat.end());
    auto pct = [&](double q) { return lat.empty() ? 0.f : lat[std::min(lat.size() - 1, (size_t)(q * lat.size()))]; };
    printf("{\"workload\": \"conference %dx%d QP32 veryfast paced at %.0f fps, nothing in flight, encode%s per stream\", \"streams\": %d, "
           "\"frames_per_stream\": %d, \"achieved_fps_per_stream\": %.2f, \"latency_ms\": {\"p50\": %.2f, \"p95\": %.2f, \"p99\": %.2f, \"max\": %.2f}, "
           "\"pictures_timed\": %zu, \"all_pictures_decoded\": %s, \"gpus\": %d, \"streams_hash\": \"%016llx\"}\n",
           W, H, PACE_FPS, ENC_ONLY ? "" : "+decode", STREAMS, FRAMES, fps / STREAMS, pct(0.50), pct(0.95), pct(0.99), lat.empty() ? 0.f : lat.back(),
           lat.size(), all ? "true" : "false", NDEV, (unsigned long long)all_hash);
    return lat.empty() || !all ? 1 : 0;
  }
  printf("{\"workload\": \"conference %dx%d QP32 veryfast, encode%s per stream, C++ harness over the C ABI\", \"streams\": %d, "
         "\"frames_per_stream\": %d, \"decoder_frame_threads\": %d, \"seconds\": %.3f, \"aggregate_fps\": %.1f, "
         "\"fps_per_stream\": %.1f, \"streams_sustained_at_30fps\": %d, \"all_pictures_decoded\": %s, \"gpus\": %d, \"streams_hash\": \"%016llx\"}\n",
         W, H, ENC_ONLY ? "" : "+decode", STREAMS, FRAMES, DEC_THREADS, dt, fps, fps / STREAMS, (int)(fps / 30), all ? "true" : "false", NDEV,
         (unsigned long long)all_hash);
  return all ? 0 : 1;
}
