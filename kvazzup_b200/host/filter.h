// Qt-free restatement of the reference's Filter / Data runtime, header only.
//
// Same names, ownership and queueing rules as /root/reference/src/media/processing/filter.h:27-261 and
// filter.cpp:151-222 (putInput: bounded deque, overflow policy), :297-340 (getInput), :364-417
// (sendOutput: every consumer but the last gets a deep copy, the last one the original), :425-443 (run:
// one thread per filter, sleeping while there is no input), :485-499 (deepDataCopy), :516-532
// (isHEVCIntra / isHEVCInter) -- with std::thread / std::mutex / std::condition_variable in place of
// QThread / QMutex / QWaitCondition, and without the statistics and logging hooks.  The GPU filters of
// kvazzup_b200/host/filters.h derive from this class exactly as the reference's filters derive from its
// Filter; with Qt present the same bodies compile against the real base class.
#pragma once
#include <stdint.h>
#include <string.h>

#include <atomic>
#include <condition_variable>
#include <deque>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace b200host {

// filter.h:27-52 (bit-flag values are the reference's)
enum DataType {DT_NONE = 0, DT_YUV420VIDEO = 1, DT_YUV422VIDEO = 1 << 1, DT_NV12VIDEO = 1 << 2, DT_NV21VIDEO = 1 << 3,
               DT_YUYVVIDEO = 1 << 4, DT_UYVYVIDEO = 1 << 5, DT_ARGBVIDEO = 1 << 7, DT_BGRAVIDEO = 1 << 8, DT_ABGRVIDEO = 1 << 9,
               DT_RGB32VIDEO = 1 << 10, DT_RGB24VIDEO = 1 << 11, DT_BGRXVIDEO = 1 << 12, DT_MJPEGVIDEO = 1 << 13,
               DT_HEVCVIDEO = 1 << 14, DT_RAWAUDIO = 1 << 15, DT_OPUSAUDIO = 1 << 16};
enum DataSource {DS_UNKNOWN, DS_LOCAL, DS_REMOTE};
enum HEVC_NAL_UNIT_TYPE {TRAIL_R = 1, IDR_W_RADL = 19, VPS_NUT = 32, SPS_NUT = 33, PPS_NUT = 34};

// global.h:55-59
struct RoiMap {
  int width = 0, height = 0;
  std::unique_ptr<int8_t[]> data;
};

// filter.h:59-70
struct VideoInfo {
  int16_t width = 0, height = 0;
  int32_t framerateNumerator = 0, framerateDenominator = 0;
  bool flippedVertically = false, flippedHorizontally = false;
  RoiMap roi;
};

// filter.h:77-92
struct Data {
  DataSource source = DS_UNKNOWN;
  DataType type = DT_NONE;
  std::unique_ptr<uint8_t[]> data;
  uint32_t data_size = 0;
  int64_t creationTimestamp = -1;
  int64_t presentationTimestamp = -1;
  std::unique_ptr<VideoInfo> vInfo;
};

class Filter {
 public:
  // filter.cpp:45-68.  The reference sizes the input buffer per kind of filter (maxBufferSize_); -1 = unbounded.
  Filter(std::string id, std::string name, DataType input, DataType output, int maxBufferSize = 10)
      : id_(std::move(id)), name_(std::move(name)), input_(input), output_(output), maxBufferSize_(maxBufferSize) {}
  virtual ~Filter() { stop(); if (thread_.joinable()) thread_.join(); }

  virtual bool init() { return true; }                 // false: the graph drops the filter (filtergraph.cpp:471-479)
  virtual void updateSettings() {}

  void addOutConnection(std::shared_ptr<Filter> out) { std::lock_guard<std::mutex> l(connectionMutex_); outConnections_.push_back(std::move(out)); }
  void addDataOutCallback(std::function<void(std::unique_ptr<Data>)> cb) { std::lock_guard<std::mutex> l(connectionMutex_); outDataCallbacks_.push_back(std::move(cb)); }

  // filter.cpp:168-221: called by the producer's thread
  void putInput(std::unique_ptr<Data> data)
  {
    if (!data) return;
    ++inputTaken_;
    std::lock_guard<std::mutex> l(bufferMutex_);
    inBuffer_.push_back(std::move(data));
    if (maxBufferSize_ != -1 && inBuffer_.size() >= (size_t)maxBufferSize_) {
      if (inBuffer_[0]->type == DT_HEVCVIDEO) {
        // The reference means to discard up to the next intra picture; its loop looks for the first
        // buffer that is NOT an intra NAL and drops everything before it (:179-197).  Restated as
        // written, since that is what a drop-in must reproduce.
        for (size_t i = 0; i < inBuffer_.size(); ++i) {
          if (!isHEVCIntra(inBuffer_[i]->data.get())) {
            for (size_t j = i; j != 0; --j) inBuffer_.pop_front();
            break;
          }
        }
      } else {
        inBuffer_.pop_front();                         // discard the oldest
      }
      ++inputDiscarded_;
    }
    hasInput_.notify_one();
  }

  void start() { running_ = true; thread_ = std::thread([this] { run(); }); }
  void stop() { running_ = false; hasInput_.notify_all(); }
  uint32_t inputTaken() const { return inputTaken_; }
  uint32_t inputDiscarded() const { return inputDiscarded_; }
  size_t buffered() { std::lock_guard<std::mutex> l(bufferMutex_); return inBuffer_.size(); }
  const std::string &name() const { return name_; }
  DataType inputType() const { return input_; }
  DataType outputType() const { return output_; }

  // filter.cpp:516-532
  static bool isHEVCIntra(const uint8_t *b) { return b[0] == 0 && b[1] == 0 && b[2] == 0 && b[3] == 1 && (b[4] >> 1) == IDR_W_RADL; }
  static bool isHEVCInter(const uint8_t *b) { return b[0] == 0 && b[1] == 0 && b[2] == 0 && b[3] == 1 && (b[4] >> 1) == TRAIL_R; }

 protected:
  virtual void process() = 0;                          // loop: while (auto in = getInput()) { ...; sendOutput(std::move(out)); }

  // filter.cpp:297-340
  std::unique_ptr<Data> getInput()
  {
    std::lock_guard<std::mutex> l(bufferMutex_);
    std::unique_ptr<Data> r;
    if (!inBuffer_.empty()) { r = std::move(inBuffer_.front()); inBuffer_.pop_front(); }
    return r;
  }

  // filter.cpp:364-417: all consumers but the last one get a deep copy
  void sendOutput(std::unique_ptr<Data> output)
  {
    if (!output) return;
    std::lock_guard<std::mutex> l(connectionMutex_);
    if (outDataCallbacks_.empty() && outConnections_.empty()) return;
    if (!outDataCallbacks_.empty()) {
      for (size_t i = 0; i + 1 < outDataCallbacks_.size(); ++i) outDataCallbacks_[i](deepDataCopy(output.get()));
      if (!outConnections_.empty()) outDataCallbacks_.back()(deepDataCopy(output.get()));
      else { outDataCallbacks_.back()(std::move(output)); return; }
    }
    for (size_t i = 0; i + 1 < outConnections_.size(); ++i) outConnections_[i]->putInput(deepDataCopy(output.get()));
    outConnections_.back()->putInput(std::move(output));
  }

  // filter.cpp:222-258
  static std::unique_ptr<Data> initializeData(DataType type, DataSource source)
  {
    std::unique_ptr<Data> d(new Data);
    d->type = type; d->source = source; d->data_size = 0; d->creationTimestamp = 0; d->presentationTimestamp = 0;
    d->vInfo.reset(new VideoInfo);
    return d;
  }

  // filter.cpp:455-499
  static std::unique_ptr<Data> deepDataCopy(const Data *o)
  {
    std::unique_ptr<Data> c(new Data);
    c->source = o->source; c->type = o->type; c->data_size = o->data_size;
    c->creationTimestamp = o->creationTimestamp; c->presentationTimestamp = o->presentationTimestamp;
    if (o->vInfo) {
      c->vInfo.reset(new VideoInfo);
      c->vInfo->width = o->vInfo->width; c->vInfo->height = o->vInfo->height;
      c->vInfo->framerateNumerator = o->vInfo->framerateNumerator; c->vInfo->framerateDenominator = o->vInfo->framerateDenominator;
      c->vInfo->flippedVertically = o->vInfo->flippedVertically; c->vInfo->flippedHorizontally = o->vInfo->flippedHorizontally;
      if (o->vInfo->roi.data) {
        const size_t n = (size_t)o->vInfo->roi.width * o->vInfo->roi.height;
        c->vInfo->roi.width = o->vInfo->roi.width; c->vInfo->roi.height = o->vInfo->roi.height;
        c->vInfo->roi.data.reset(new int8_t[n]);
        memcpy(c->vInfo->roi.data.get(), o->vInfo->roi.data.get(), n);
      }
    }
    c->data.reset(new uint8_t[o->data_size]);
    memcpy(c->data.get(), o->data.get(), o->data_size);
    return c;
  }

  std::string id_, name_;
  DataType input_, output_;

 private:
  // filter.cpp:425-443
  void run()
  {
    while (running_) {
      {
        std::unique_lock<std::mutex> l(bufferMutex_);
        hasInput_.wait(l, [this] { return !running_ || !inBuffer_.empty(); });
      }
      if (!running_) break;
      process();
    }
  }

  int maxBufferSize_;
  std::mutex bufferMutex_, connectionMutex_;
  std::condition_variable hasInput_;
  std::deque<std::unique_ptr<Data>> inBuffer_;
  std::vector<std::shared_ptr<Filter>> outConnections_;
  std::vector<std::function<void(std::unique_ptr<Data>)>> outDataCallbacks_;
  std::atomic<bool> running_{false};
  std::atomic<uint32_t> inputTaken_{0}, inputDiscarded_{0};
  std::thread thread_;
};

}  // namespace b200host
