// The four media filters of the hot path as Filter subclasses over the C ABI of libb200media.so
// (8f-2).  Each process() body follows the reference filter it stands
// for, line by line where the logic is the filter's own:
//   LibYUVConverter   src/media/processing/libyuvconverter.cpp:20-136
//   KvazaarFilter     src/media/processing/kvazaarfilter.cpp:122-311 (init), :374-450 (feedInput), :453-484 (parseEncodedFrame)
//   UvgRTPShim        what uvgrtpsender.cpp:89-118 / uvgrtpreceiver.cpp:54-116 do between the two: one NAL per Data,
//                     4-byte start code in front (no network: the access unit is split and forwarded)
//   OpenHEVCFilter    src/media/processing/openhevcfilter.cpp:28-99 (init), :103-189 (process), :192-239 (sendDecodedOutput)
//   YUVtoRGB32        src/media/processing/yuvtorgb32.cpp:29-64
// With Qt present the same bodies compile against the reference's own Filter base.
#pragma once
#include <map>
#include <string>

#include "../../include/b200_kvazaar.h"
#include "../../include/b200_openhevc.h"
#include "../../include/b200_rtp.h"
#include "../../include/b200media.h"
#include "filter.h"

namespace b200host {

// the uvgComm.ini keys the filters read (settingskeys.h); values as strings like QSettings hands them out
using Settings = std::map<std::string, std::string>;
inline std::string setting(const Settings &s, const char *key, const char *dflt)
{
  auto it = s.find(key);
  return it == s.end() ? dflt : it->second;
}

class LibYUVConverter : public Filter {
 public:
  LibYUVConverter(std::string id, DataType input) : Filter(std::move(id), "libyuv", input, DT_YUV420VIDEO) {}

 protected:
  void process() override
  {
    std::unique_ptr<Data> input = getInput();
    while (input) {
      if (input->type == DT_YUV420VIDEO) {             // :30-35 forwarded untouched
        sendOutput(std::move(input));
        input = getInput();
        continue;
      }
      uint32_t fourcc = 0;
      switch (input->type) {                            // :39-94
      case DT_YUV422VIDEO: fourcc = B200_FOURCC_I422; break;
      case DT_NV12VIDEO: fourcc = B200_FOURCC_NV12; break;
      case DT_NV21VIDEO: fourcc = B200_FOURCC_NV21; break;
      case DT_YUYVVIDEO: fourcc = B200_FOURCC_YUYV; break;
      case DT_UYVYVIDEO: fourcc = B200_FOURCC_UYVY; break;
      case DT_ARGBVIDEO: fourcc = B200_FOURCC_ARGB; break;
      case DT_BGRAVIDEO: fourcc = B200_FOURCC_BGRA; break;
      case DT_ABGRVIDEO: fourcc = B200_FOURCC_ABGR; break;
      case DT_RGB32VIDEO: fourcc = B200_FOURCC_RGBA; break;
      case DT_BGRXVIDEO: fourcc = B200_FOURCC_24BG; break;
      case DT_MJPEGVIDEO: fourcc = B200_FOURCC_MJPG; break;
      default: break;
      }
      const int w = input->vInfo->width, h = input->vInfo->height;
      const uint32_t out_size = (uint32_t)(w * h + w * h / 2);                        // :106-110
      std::unique_ptr<uint8_t[]> yuv(new uint8_t[out_size]);
      uint8_t *y = yuv.get(), *u = y + w * h, *v = u + w * h / 4;
      b200_ConvertToI420(input->data.get(), input->data_size, y, w, u, (w + 1) / 2, v, (w + 1) / 2, 0, 0, w, h, w, h, 0, fourcc);   // :120-127, return value ignored
      input->type = DT_YUV420VIDEO;
      input->data = std::move(yuv);
      input->data_size = out_size;
      sendOutput(std::move(input));
      input = getInput();
    }
  }
};

class KvazaarFilter : public Filter {
 public:
  KvazaarFilter(std::string id, Settings s) : Filter(std::move(id), "Kvazaar", DT_YUV420VIDEO, DT_HEVCVIDEO), settings_(std::move(s)) {}
  ~KvazaarFilter() override { stop(); close(); }

  bool init() override                                 // :122-311
  {
    api_ = kvz_api_get(8);
    if (!api_) return false;
    config_ = api_->config_alloc();
    if (!config_) return false;
    api_->config_init(config_);
    auto parse = [&](const char *name, const std::string &value) { return api_->config_parse(config_, name, value.c_str()) == 1; };
    parse("preset", setting(settings_, "video/Preset", "ultrafast"));
    parse("input-res", setting(settings_, "video/ResolutionWidth", "640") + "x" + setting(settings_, "video/ResolutionHeight", "480"));
    parse("input-fps", setting(settings_, "video/FramerateNumerator", "30") + "/" + setting(settings_, "video/FramerateDenominator", "1"));
    parse("owf", setting(settings_, "video/OWF", "0"));
    parse("wpp", setting(settings_, "video/WPP", "1"));
    parse("qp", setting(settings_, "video/QP", "32"));
    parse("period", setting(settings_, "video/Intra", "64"));
    parse("vps-period", setting(settings_, "video/VPS", "1"));
    config_->target_bitrate = atoi(setting(settings_, "video/bitrate", "0").c_str());
    parse("gop", "lp-g4d3t1");
    config_->hash = KVZ_HASH_NONE;
    enc_ = api_->encoder_open(config_);
    if (!enc_) return false;
    for (int i = 0; i < config_->owf + 1; ++i) {        // :299 createInputVector(owf + 1)
      kvz_picture *p = api_->picture_alloc(config_->width, config_->height);
      if (!p) return false;
      inputPics_.push_back(p);
    }
    return true;
  }

 protected:
  void process() override
  {
    std::unique_ptr<Data> input = getInput();
    while (input) {
      feedInput(std::move(input));
      input = getInput();
    }
  }

 private:
  void close()
  {
    if (!api_) return;
    for (kvz_picture *p : inputPics_) api_->picture_free(p);
    inputPics_.clear();
    if (enc_) api_->encoder_close(enc_);
    if (config_) api_->config_destroy(config_);
    enc_ = nullptr; config_ = nullptr;
  }

  void feedInput(std::unique_ptr<Data> input)          // :374-450
  {
    if (input->vInfo->width != config_->width || input->vInfo->height != config_->height) return;   // (the reference re-initialises: :381-404)
    kvz_picture *pic = inputPics_[nextInputPic_];
    nextInputPic_ = (nextInputPic_ + 1) % (int)inputPics_.size();
    const size_t ysz = (size_t)config_->width * config_->height;
    memcpy(pic->y, input->data.get(), ysz);                                            // :410-418
    memcpy(pic->u, input->data.get() + ysz, ysz / 4);
    memcpy(pic->v, input->data.get() + ysz + ysz / 4, ysz / 4);
    pic->pts = input->presentationTimestamp;                                           // :420
    pic->roi.width = pic->roi.height = 0; pic->roi.roi_array = nullptr;
    encodingFrames_.push_front(std::move(input));                                      // :433
    kvz_data_chunk *data_out = nullptr;
    uint32_t len_out = 0;
    kvz_picture *recon = nullptr;
    kvz_frame_info info;
    api_->encoder_encode(enc_, pic, &data_out, &len_out, &recon, nullptr, &info);     // :435
    while (data_out != nullptr) {                                                      // :440-449
      parseEncodedFrame(data_out, len_out, recon);
      api_->encoder_encode(enc_, nullptr, &data_out, &len_out, &recon, nullptr, &info);
    }
  }

  void parseEncodedFrame(kvz_data_chunk *data_out, uint32_t len_out, kvz_picture *recon)   // :453-484
  {
    std::unique_ptr<Data> encoded = std::move(encodingFrames_.back());
    encodingFrames_.pop_back();
    std::unique_ptr<uint8_t[]> frame(new uint8_t[len_out]);
    uint8_t *w = frame.get();
    for (kvz_data_chunk *c = data_out; c != nullptr; c = c->next) { memcpy(w, c->data, c->len); w += c->len; }   // :465-474
    api_->chunk_free(data_out);
    api_->picture_free(recon);
    encoded->type = DT_HEVCVIDEO;
    encoded->data_size = len_out;
    encoded->data = std::move(frame);
    sendOutput(std::move(encoded));
  }

  Settings settings_;
  const kvz_api *api_ = nullptr;
  kvz_config *config_ = nullptr;
  kvz_encoder *enc_ = nullptr;
  std::vector<kvz_picture *> inputPics_;
  int nextInputPic_ = 0;
  std::deque<std::unique_ptr<Data>> encodingFrames_;
};

// Stand-in for the uvgRTP sender + receiver pair: the access unit leaves as NAL units, each of which
// arrives as its own Data with a 4-byte start code (uvgrtpreceiver.cpp:87-111).
class UvgRTPShim : public Filter {
 public:
  explicit UvgRTPShim(std::string id) : Filter(std::move(id), "RTP shim", DT_HEVCVIDEO, DT_HEVCVIDEO, -1) {}

 protected:
  void process() override
  {
    std::unique_ptr<Data> input = getInput();
    while (input) {
      b200_nal_span spans[64];
      const int n = b200_annexb_split(input->data.get(), input->data_size, spans, 64);
      for (int i = 0; i < n && i < 64; i++) {
        std::unique_ptr<Data> nal = deepDataCopy(input.get());
        nal->source = DS_REMOTE;
        nal->data_size = (uint32_t)(4 + spans[i].length);
        nal->data.reset(new uint8_t[nal->data_size]);
        nal->data[0] = nal->data[1] = nal->data[2] = 0; nal->data[3] = 1;
        memcpy(nal->data.get() + 4, input->data.get() + spans[i].offset, spans[i].length);
        sendOutput(std::move(nal));
      }
      input = getInput();
    }
  }
};

class OpenHEVCFilter : public Filter {
 public:
  OpenHEVCFilter(std::string id, int threads = 1, int threadType = 2)
      : Filter(std::move(id), "OpenHEVC", DT_HEVCVIDEO, DT_YUV420VIDEO, -1), threads_(threads), threadType_(threadType) {}
  ~OpenHEVCFilter() override { stop(); if (handle_) { libOpenHevcFlush(handle_); libOpenHevcClose(handle_); } }

  bool init() override                                 // :28-99
  {
    handle_ = libOpenHevcInit(threads_, threadType_);
    if (libOpenHevcStartDecoder(handle_) == -1) return false;
    libOpenHevcSetTemporalLayer_id(handle_, 0);
    libOpenHevcSetActiveDecoders(handle_, 0);
    libOpenHevcSetViewLayers(handle_, 0);
    return true;
  }
  uint32_t discardedFrames() const { return discardedFrames_; }

 protected:
  void process() override                              // :103-189
  {
    std::unique_ptr<Data> input = getInput();
    while (input) {
      const uint8_t *buff = input->data.get();
      const uint8_t nalType = buff[4] >> 1;
      if (nalType == VPS_NUT) vpsReceived_ = true;
      if (nalType == SPS_NUT) spsReceived_ = true;
      if (nalType == PPS_NUT) ppsReceived_ = true;
      const bool vcl = nalType <= 31;
      if ((vpsReceived_ && spsReceived_ && ppsReceived_) || !vcl) {
        int gotPicture = libOpenHevcDecode(handle_, input->data.get(), (int)input->data_size, input->presentationTimestamp);
        if (vcl) decodingFrames_.push_front(std::move(input));
        if (gotPicture > 0) sendDecodedOutput(gotPicture);      // <= -1: "Error while decoding!", 0: nothing yet
      } else {
        ++discardedFrames_;
      }
      input = getInput();
    }
  }

 private:
  void sendDecodedOutput(int &gotPicture)              // :192-239
  {
    OpenHevc_Frame f;
    if ((gotPicture = libOpenHevcGetOutput(handle_, gotPicture, &f)) > 0) {
      std::unique_ptr<Data> decoded = std::move(decodingFrames_.back());
      decodingFrames_.pop_back();
      libOpenHevcGetPictureInfo(handle_, &f.frameInfo);
      decoded->vInfo->width = (int16_t)f.frameInfo.nWidth;
      decoded->vInfo->height = (int16_t)f.frameInfo.nHeight;
      const int w = decoded->vInfo->width, h = decoded->vInfo->height;
      const uint32_t size = (uint32_t)(w * h + w * h / 2);
      std::unique_ptr<uint8_t[]> yuv(new uint8_t[size]);
      uint8_t *pY = yuv.get(), *pU = pY + w * h, *pV = pU + w * h / 4;
      const uint32_t s_stride = f.frameInfo.nYPitch, qs_stride = f.frameInfo.nUPitch / 2;
      for (int i = 0; i < h; i++) {
        memcpy(pY, (uint8_t *)f.pvY + i * s_stride, w); pY += w;
        if (!(i % 2)) {
          memcpy(pU, (uint8_t *)f.pvU + i * qs_stride, w / 2); pU += w / 2;
          memcpy(pV, (uint8_t *)f.pvV + i * qs_stride, w / 2); pV += w / 2;
        }
      }
      decoded->type = DT_YUV420VIDEO;
      decoded->vInfo->framerateNumerator = f.frameInfo.frameRate.num;
      decoded->vInfo->framerateDenominator = f.frameInfo.frameRate.den;
      decoded->data_size = size;
      decoded->data = std::move(yuv);
      sendOutput(std::move(decoded));
    }
  }

  int threads_, threadType_;
  OpenHevc_Handle handle_ = nullptr;
  bool vpsReceived_ = false, spsReceived_ = false, ppsReceived_ = false;
  uint32_t discardedFrames_ = 0;
  std::deque<std::unique_ptr<Data>> decodingFrames_;
};

class YUVtoRGB32 : public Filter {
 public:
  explicit YUVtoRGB32(std::string id) : Filter(std::move(id), "YUVtoRGB32", DT_YUV420VIDEO, DT_RGB32VIDEO) {}

 protected:
  void process() override                              // yuvtorgb32.cpp:29-64
  {
    std::unique_ptr<Data> input = getInput();
    while (input) {
      const int w = input->vInfo->width, h = input->vInfo->height;
      const uint32_t size = (uint32_t)(w * h * 4);
      std::unique_ptr<uint8_t[]> rgb(new uint8_t[size]);
      b200_yuv420_to_rgb32(input->data.get(), rgb.get(), (uint16_t)w, (uint16_t)h);
      input->type = DT_RGB32VIDEO;
      input->data = std::move(rgb);
      input->data_size = size;
      sendOutput(std::move(input));
      input = getInput();
    }
  }
};

}  // namespace b200host
