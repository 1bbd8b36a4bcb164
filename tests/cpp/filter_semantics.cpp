// CPU-only checks of the Filter runtime restatement (kvazzup_b200/host/filter.h) against the rules of
// the reference's filter.cpp: bounded queue + drop policy (:168-221), fan-out deep copy (:364-417),
// one thread per filter (:425-443).  No GPU, no library: dummy filters.
#include <stdio.h>

#include <chrono>

#include "../../kvazzup_b200/host/filter.h"

using namespace b200host;

#define CHECK(c) do { if (!(c)) { printf("FAIL line %d: %s\n", __LINE__, #c); return 1; } } while (0)

static std::unique_ptr<Data> make(DataType t, uint8_t fill, int nal_type = -1)
{
  std::unique_ptr<Data> d(new Data);
  d->type = t; d->data_size = 16; d->data.reset(new uint8_t[16]);
  memset(d->data.get(), fill, 16);
  if (nal_type >= 0) { d->data[0] = d->data[1] = d->data[2] = 0; d->data[3] = 1; d->data[4] = (uint8_t)(nal_type << 1); }
  d->vInfo.reset(new VideoInfo);
  d->vInfo->width = 4; d->vInfo->height = 4;
  return d;
}

struct Passthrough : Filter {
  Passthrough(const char *n, int cap) : Filter(n, n, DT_YUV420VIDEO, DT_YUV420VIDEO, cap) {}
  std::atomic<int> seen{0};
  void process() override { while (auto in = getInput()) { in->data[0]++; seen++; sendOutput(std::move(in)); } }
};
struct Blocked : Filter {      // never started: its queue only fills
  Blocked(DataType t, int cap) : Filter("blocked", "blocked", t, t, cap) {}
  void process() override {}
  std::unique_ptr<Data> pop() { return getInput(); }
};

int main()
{
  // drop policy, raw video: the oldest goes when the queue reaches its size
  {
    Blocked b(DT_YUV420VIDEO, 3);
    for (int i = 0; i < 5; i++) b.putInput(make(DT_YUV420VIDEO, (uint8_t)i));
    CHECK(b.buffered() == 2 && b.inputDiscarded() == 3 && b.inputTaken() == 5);
    CHECK(b.pop()->data[5] == 3 && b.pop()->data[5] == 4);
  }
  // drop policy, HEVC: everything in front of the first NAL that is not an intra NAL goes (as the reference is written)
  {
    Blocked b(DT_HEVCVIDEO, 4);
    b.putInput(make(DT_HEVCVIDEO, 0, 19));
    b.putInput(make(DT_HEVCVIDEO, 0, 19));
    b.putInput(make(DT_HEVCVIDEO, 0, 1));
    CHECK(b.buffered() == 3);
    b.putInput(make(DT_HEVCVIDEO, 0, 1));              // size reaches 4: the two intra NALs in front are dropped
    CHECK(b.buffered() == 2 && b.inputDiscarded() == 1);
    CHECK(Filter::isHEVCInter(b.pop()->data.get()));
    CHECK(Filter::isHEVCIntra(make(DT_HEVCVIDEO, 0, 19)->data.get()) && !Filter::isHEVCIntra(make(DT_HEVCVIDEO, 0, 1)->data.get()));
  }
  // fan-out: the first consumers get deep copies, the last one the original; each filter runs on its own thread
  {
    auto src = std::make_shared<Passthrough>("src", 10);
    auto a = std::make_shared<Passthrough>("a", 10), b = std::make_shared<Passthrough>("b", 10);
    std::mutex m;
    std::vector<std::unique_ptr<Data>> got_a, got_b;
    src->addOutConnection(a);
    src->addOutConnection(b);
    a->addDataOutCallback([&](std::unique_ptr<Data> d) { std::lock_guard<std::mutex> l(m); got_a.push_back(std::move(d)); });
    b->addDataOutCallback([&](std::unique_ptr<Data> d) { std::lock_guard<std::mutex> l(m); got_b.push_back(std::move(d)); });
    src->start(); a->start(); b->start();
    for (int i = 0; i < 8; i++) src->putInput(make(DT_YUV420VIDEO, 10));
    for (int spin = 0; spin < 2000; spin++) {
      { std::lock_guard<std::mutex> l(m); if (got_a.size() == 8 && got_b.size() == 8) break; }
      std::this_thread::sleep_for(std::chrono::milliseconds(1));
    }
    src->stop(); a->stop(); b->stop();
    CHECK(got_a.size() == 8 && got_b.size() == 8);
    for (int i = 0; i < 8; i++) {
      CHECK(got_a[i]->data.get() != got_b[i]->data.get());       // different buffers: one is the deep copy
      CHECK(got_a[i]->data[0] == 12 && got_b[i]->data[0] == 12); // src + own increment each, no cross-talk
      CHECK(got_a[i]->vInfo && got_a[i]->vInfo->width == 4);
    }
    CHECK(src->seen == 8 && a->seen == 8 && b->seen == 8);
  }
  printf("OK\n");
  return 0;
}
