// BASELINE config 1, "uvgComm loopback media pipeline", as a filter graph of the reference's shape:
//
//   camera (YUYV frames) -> LibYUVConverter -> KvazaarFilter -> RTP shim (one NAL per Data) -> OpenHEVCFilter -> YUVtoRGB32 -> display
//                                   \-> self view: YUVtoRGB32 (second consumer of the I420 frame: gets the deep copy, filter.cpp:364-417)
//
// every filter on its own thread behind a bounded input queue (kvazzup_b200/host/filter.h), all
// arithmetic in libb200media.so.  Prints one JSON line: pictures in / displayed, FNV-1a hashes of the
// displayed RGB32 pictures and of the self-view pictures, frames per second.
//
//   loopback_pipeline <frames.yuyv> <w> <h> <frames_in_file> <frames> [preset] [qp] [pace_fps]
#include <stdio.h>
#include <stdlib.h>

#include <chrono>
#include <thread>

#include "../kvazzup_b200/host/filters.h"

using namespace b200host;

struct Sink {
  std::mutex m;
  uint64_t hash = 1469598103934665603ull;
  int pictures = 0;
  std::chrono::steady_clock::time_point last;          // when the newest picture arrived
  FILE *dump = nullptr;                                // B200_LOOPBACK_DUMP: the pictures themselves, for the parity test
  void take(std::unique_ptr<Data> d)
  {
    std::lock_guard<std::mutex> l(m);
    for (uint32_t i = 0; i < d->data_size; i++) { hash ^= d->data[i]; hash *= 1099511628211ull; }
    if (dump) fwrite(d->data.get(), 1, d->data_size, dump);
    pictures++;
    last = std::chrono::steady_clock::now();
  }
};

int main(int argc, char **argv)
{
  if (argc < 6) { fprintf(stderr, "usage: %s frames.yuyv w h frames_in_file frames [preset] [qp] [pace_fps]\n", argv[0]); return 2; }
  const int w = atoi(argv[2]), h = atoi(argv[3]), nfile = atoi(argv[4]), frames = atoi(argv[5]);
  const char *preset = argc > 6 ? argv[6] : "ultrafast";
  const char *qp = argc > 7 ? argv[7] : "32";
  const double pace = argc > 8 ? atof(argv[8]) : 0;
  const size_t fb = (size_t)w * h * 2;
  std::vector<uint8_t> yuyv(fb * nfile);
  FILE *f = fopen(argv[1], "rb");
  if (!f || fread(yuyv.data(), 1, yuyv.size(), f) != yuyv.size()) { fprintf(stderr, "cannot read %s\n", argv[1]); return 2; }
  fclose(f);
  if (b200_device_count() <= 0) { fprintf(stderr, "no CUDA device: %s\n", b200_last_error()); return 3; }

  Settings s = {{"video/Preset", preset}, {"video/QP", qp}, {"video/ResolutionWidth", std::to_string(w)},
                {"video/ResolutionHeight", std::to_string(h)}, {"video/Intra", "64"}, {"video/OWF", "0"}};
  auto conv = std::make_shared<LibYUVConverter>("camera-conv", DT_YUYVVIDEO);
  auto enc = std::make_shared<KvazaarFilter>("encoder", s);
  auto rtp = std::make_shared<UvgRTPShim>("rtp");
  auto dec = std::make_shared<OpenHEVCFilter>("decoder");
  auto disp = std::make_shared<YUVtoRGB32>("display-conv");
  auto self = std::make_shared<YUVtoRGB32>("selfview-conv");
  Sink display, selfview;
  if (const char *prefix = getenv("B200_LOOPBACK_DUMP")) {
    display.dump = fopen((std::string(prefix) + ".display.rgb").c_str(), "wb");
    selfview.dump = fopen((std::string(prefix) + ".selfview.rgb").c_str(), "wb");
  }
  if (!conv->init() || !enc->init() || !rtp->init() || !dec->init() || !disp->init() || !self->init()) {
    fprintf(stderr, "filter init failed: %s\n", b200_last_error());
    return 3;
  }
  conv->addOutConnection(self);                        // first consumer: deep copy
  conv->addOutConnection(enc);                         // last consumer: the frame itself
  enc->addOutConnection(rtp);
  rtp->addOutConnection(dec);
  dec->addOutConnection(disp);
  disp->addDataOutCallback([&](std::unique_ptr<Data> d) { display.take(std::move(d)); });
  self->addDataOutCallback([&](std::unique_ptr<Data> d) { selfview.take(std::move(d)); });
  for (auto &flt : std::vector<std::shared_ptr<Filter>>{conv, enc, rtp, dec, disp, self}) flt->start();

  const auto t0 = std::chrono::steady_clock::now();
  for (int t = 0; t < frames; t++) {
    if (pace > 0) std::this_thread::sleep_until(t0 + std::chrono::duration_cast<std::chrono::steady_clock::duration>(std::chrono::duration<double>(t / pace)));
    std::unique_ptr<Data> d(new Data);                 // camerafilter.cpp: a fresh buffer per frame
    d->source = DS_LOCAL; d->type = DT_YUYVVIDEO; d->data_size = (uint32_t)fb;
    d->data.reset(new uint8_t[fb]);
    memcpy(d->data.get(), &yuyv[(size_t)(t % nfile) * fb], fb);
    d->creationTimestamp = d->presentationTimestamp = t;
    d->vInfo.reset(new VideoInfo);
    d->vInfo->width = (int16_t)w; d->vInfo->height = (int16_t)h; d->vInfo->framerateNumerator = 30; d->vInfo->framerateDenominator = 1;
    // unpaced: do not outrun the bounded queues (a camera never does), or the drop policy kicks in
    while (pace <= 0 && (conv->buffered() > 4 || enc->buffered() > 4 || rtp->buffered() > 4 || dec->buffered() > 4 ||
                         disp->buffered() > 4 || self->buffered() > 4))
      std::this_thread::sleep_for(std::chrono::microseconds(50));
    conv->putInput(std::move(d));
  }
  for (int spin = 0; spin < 20000 && (display.pictures < frames || selfview.pictures < frames); spin++)
    std::this_thread::sleep_for(std::chrono::milliseconds(1));
  const double dt = std::chrono::duration<double>(std::max(display.last, selfview.last) - t0).count();
  const uint32_t dropped = conv->inputDiscarded() + enc->inputDiscarded() + dec->inputDiscarded() + disp->inputDiscarded() + self->inputDiscarded();
  for (auto &flt : std::vector<std::shared_ptr<Filter>>{conv, enc, rtp, dec, disp, self}) flt->stop();
  if (display.dump) fclose(display.dump);
  if (selfview.dump) fclose(selfview.dump);
  printf("{\"workload\": \"loopback %dx%d YUYV -> I420 -> HEVC %s QP%s -> NALs -> decode -> RGB32, filter graph\", \"frames_in\": %d, "
         "\"displayed\": %d, \"selfview\": %d, \"dropped\": %u, \"seconds\": %.3f, \"fps\": %.1f, \"display_hash\": \"%016llx\", "
         "\"selfview_hash\": \"%016llx\"}\n",
         w, h, preset, qp, frames, display.pictures, selfview.pictures, dropped, dt, frames / dt,
         (unsigned long long)display.hash, (unsigned long long)selfview.hash);
  return display.pictures == frames && selfview.pictures == frames ? 0 : 1;
}
