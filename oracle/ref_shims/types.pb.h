// types.pb.h — SHIM (test infrastructure written for this repo; not protoc output).
//
// The reference compiles types.proto with protoc (CMakeLists.txt:28); protoc and libprotobuf are
// not in this image.  This header offers the generated-code API that solve.cc:432-472,646-679
// calls, backed by a hand-written proto3 wire codec (oracle/ref_shims/types_pb_shim.cc) of
// exactly the messages of /root/reference/types.proto:
//   MatchingFile{ repeated ImagePair image_pairs = 1 }, ImagePair{ string image_name1 = 1; float
//   fact1 = 2; string image_name2 = 3; float fact2 = 4; repeated Match matches = 5 },
//   Match{ uint32 feature_idx1 = 1; uint32 feature_idx2 = 2; float similarity = 3; repeated
//   Displacement disp1 = 4; disp2 = 5 }, Displacement{ float di = 1; float dj = 2 },
//   SolutionFile{ repeated Image images = 1 }, Image{ string image_name = 1; float fact = 2;
//   repeated Displacement displacements = 3 }, Image.Displacement{ uint32 feature_idx = 1;
//   float di = 2; float dj = 3 }.
#ifndef LFR_SHIM_TYPES_PB_H_
#define LFR_SHIM_TYPES_PB_H_

// solve.cc relies on these being pulled in transitively by the headers it includes
#include <algorithm>
#include <cassert>
#include <iterator>
#include <ostream>
#include <queue>
#include <set>
#include <string>
#include <tuple>
#include <unordered_map>
#include <vector>

#include "google/protobuf/io/coded_stream.h"

#define GOOGLE_PROTOBUF_VERIFY_VERSION

class MatchingFile_ImagePair_Match_Displacement {
 public:
  MatchingFile_ImagePair_Match_Displacement() : di_(0), dj_(0) {}
  float di() const { return di_; }
  float dj() const { return dj_; }
  float di_, dj_;
};

class MatchingFile_ImagePair_Match {
 public:
  MatchingFile_ImagePair_Match() : feature_idx1_(0), feature_idx2_(0), similarity_(0) {}
  unsigned feature_idx1() const { return feature_idx1_; }
  unsigned feature_idx2() const { return feature_idx2_; }
  float similarity() const { return similarity_; }
  int disp1_size() const { return (int)disp1_.size(); }
  int disp2_size() const { return (int)disp2_.size(); }
  const MatchingFile_ImagePair_Match_Displacement& disp1(int i) const { return disp1_[i]; }
  const MatchingFile_ImagePair_Match_Displacement& disp2(int i) const { return disp2_[i]; }
  unsigned feature_idx1_, feature_idx2_;
  float similarity_;
  std::vector<MatchingFile_ImagePair_Match_Displacement> disp1_, disp2_;
};

class MatchingFile_ImagePair {
 public:
  MatchingFile_ImagePair() : fact1_(0), fact2_(0) {}
  const std::string& image_name1() const { return image_name1_; }
  const std::string& image_name2() const { return image_name2_; }
  float fact1() const { return fact1_; }
  float fact2() const { return fact2_; }
  int matches_size() const { return (int)matches_.size(); }
  const MatchingFile_ImagePair_Match& matches(int i) const { return matches_[i]; }
  std::string image_name1_, image_name2_;
  float fact1_, fact2_;
  std::vector<MatchingFile_ImagePair_Match> matches_;
};

class MatchingFile {
 public:
  bool ParseFromCodedStream(google::protobuf::io::CodedInputStream* input);
  int image_pairs_size() const { return (int)image_pairs_.size(); }
  const MatchingFile_ImagePair& image_pairs(int i) const { return image_pairs_[i]; }
  std::vector<MatchingFile_ImagePair> image_pairs_;
};

class SolutionFile_Image_Displacement {
 public:
  SolutionFile_Image_Displacement() : feature_idx_(0), di_(0), dj_(0) {}
  void set_feature_idx(unsigned v) { feature_idx_ = v; }
  void set_di(float v) { di_ = v; }
  void set_dj(float v) { dj_ = v; }
  unsigned feature_idx_;
  float di_, dj_;
};

class SolutionFile_Image {
 public:
  SolutionFile_Image() : fact_(0) {}
  void set_image_name(const std::string& s) { image_name_ = s; }
  void set_fact(float v) { fact_ = v; }
  SolutionFile_Image_Displacement* add_displacements() {
    displacements_.push_back(SolutionFile_Image_Displacement());
    return &displacements_.back();
  }
  std::string image_name_;
  float fact_;
  std::vector<SolutionFile_Image_Displacement> displacements_;
};

class SolutionFile {
 public:
  SolutionFile_Image* mutable_images(int i) { return &images_[i]; }
  SolutionFile_Image* add_images() {
    images_.push_back(SolutionFile_Image());
    return &images_.back();
  }
  int images_size() const { return (int)images_.size(); }
  bool SerializeToOstream(std::ostream* output) const;
  std::vector<SolutionFile_Image> images_;
};

#endif  // LFR_SHIM_TYPES_PB_H_
