// types_pb_shim.cc — proto3 wire codec behind oracle/ref_shims/types.pb.h (test infrastructure).
// Decoder: unknown fields are skipped, absent fields keep their zero defaults, packed and
// unpacked encodings of the scalar fields are both accepted (none is repeated here).
// Encoder: fields in field-number order, zero-valued scalars / empty strings omitted, as
// libprotobuf's proto3 serializer does.
#include <cstring>

#include "types.pb.h"

namespace {

struct Reader {
  const unsigned char* p;
  const unsigned char* end;
  bool ok;
  Reader(const unsigned char* b, size_t n) : p(b), end(b + n), ok(true) {}
  bool done() const { return p >= end; }
  unsigned long long varint() {
    unsigned long long v = 0;
    int shift = 0;
    while (p < end && shift < 64) {
      const unsigned char b = *p++;
      v |= (unsigned long long)(b & 0x7f) << shift;
      if (!(b & 0x80)) return v;
      shift += 7;
    }
    ok = false;
    return 0;
  }
  float f32() {
    if (end - p < 4) {
      ok = false;
      return 0;
    }
    float f;
    std::memcpy(&f, p, 4);
    p += 4;
    return f;
  }
  Reader sub() {
    const unsigned long long n = varint();
    if (!ok || (unsigned long long)(end - p) < n) {
      ok = false;
      return Reader(p, 0);
    }
    Reader r(p, (size_t)n);
    p += n;
    return r;
  }
  void skip(unsigned wire) {
    switch (wire) {
      case 0: varint(); break;
      case 1: if (end - p < 8) ok = false; else p += 8; break;
      case 2: sub(); break;
      case 5: if (end - p < 4) ok = false; else p += 4; break;
      default: ok = false;
    }
  }
};

bool parse_disp(Reader r, MatchingFile_ImagePair_Match_Displacement* d) {
  while (r.ok && !r.done()) {
    const unsigned long long key = r.varint();
    const unsigned field = (unsigned)(key >> 3), wire = (unsigned)(key & 7);
    if (field == 1 && wire == 5) d->di_ = r.f32();
    else if (field == 2 && wire == 5) d->dj_ = r.f32();
    else r.skip(wire);
  }
  return r.ok;
}

bool parse_match(Reader r, MatchingFile_ImagePair_Match* m) {
  while (r.ok && !r.done()) {
    const unsigned long long key = r.varint();
    const unsigned field = (unsigned)(key >> 3), wire = (unsigned)(key & 7);
    if (field == 1 && wire == 0) m->feature_idx1_ = (unsigned)r.varint();
    else if (field == 2 && wire == 0) m->feature_idx2_ = (unsigned)r.varint();
    else if (field == 3 && wire == 5) m->similarity_ = r.f32();
    else if ((field == 4 || field == 5) && wire == 2) {
      MatchingFile_ImagePair_Match_Displacement d;
      Reader s = r.sub();
      if (!r.ok || !parse_disp(s, &d)) return false;
      (field == 4 ? m->disp1_ : m->disp2_).push_back(d);
    } else r.skip(wire);
  }
  return r.ok;
}

bool parse_pair(Reader r, MatchingFile_ImagePair* p) {
  while (r.ok && !r.done()) {
    const unsigned long long key = r.varint();
    const unsigned field = (unsigned)(key >> 3), wire = (unsigned)(key & 7);
    if ((field == 1 || field == 3) && wire == 2) {
      Reader s = r.sub();
      if (!r.ok) return false;
      (field == 1 ? p->image_name1_ : p->image_name2_).assign((const char*)s.p, (size_t)(s.end - s.p));
    } else if (field == 2 && wire == 5) p->fact1_ = r.f32();
    else if (field == 4 && wire == 5) p->fact2_ = r.f32();
    else if (field == 5 && wire == 2) {
      p->matches_.push_back(MatchingFile_ImagePair_Match());
      Reader s = r.sub();
      if (!r.ok || !parse_match(s, &p->matches_.back())) return false;
    } else r.skip(wire);
  }
  return r.ok;
}

void put_varint(std::string* o, unsigned long long v) {
  while (v >= 0x80) {
    o->push_back((char)((v & 0x7f) | 0x80));
    v >>= 7;
  }
  o->push_back((char)v);
}
void put_f32(std::string* o, unsigned tag, float f) {
  if (f <= 0 && f >= 0) return;  // proto3: default value (+-0) not serialized; NaN is
  o->push_back((char)tag);
  char b[4];
  std::memcpy(b, &f, 4);
  o->append(b, 4);
}

}  // namespace

bool MatchingFile::ParseFromCodedStream(google::protobuf::io::CodedInputStream* input) {
  image_pairs_.clear();
  const unsigned char* data = nullptr;
  size_t size = 0;
  if (!input->input()->ReadAll(&data, &size)) return false;
  if ((long long)size > input->limit()) return false;
  Reader r(data, size);
  while (r.ok && !r.done()) {
    const unsigned long long key = r.varint();
    const unsigned field = (unsigned)(key >> 3), wire = (unsigned)(key & 7);
    if (field == 1 && wire == 2) {
      image_pairs_.push_back(MatchingFile_ImagePair());
      Reader s = r.sub();
      if (!r.ok || !parse_pair(s, &image_pairs_.back())) return false;
    } else {
      r.skip(wire);
    }
  }
  return r.ok;
}

bool SolutionFile::SerializeToOstream(std::ostream* output) const {
  std::string out;
  for (size_t i = 0; i < images_.size(); ++i) {
    const SolutionFile_Image& im = images_[i];
    std::string body;
    if (!im.image_name_.empty()) {
      body.push_back((char)0x0a);
      put_varint(&body, im.image_name_.size());
      body += im.image_name_;
    }
    put_f32(&body, 0x15, im.fact_);
    for (size_t k = 0; k < im.displacements_.size(); ++k) {
      const SolutionFile_Image_Displacement& d = im.displacements_[k];
      std::string db;
      if (d.feature_idx_) {
        db.push_back((char)0x08);
        put_varint(&db, d.feature_idx_);
      }
      put_f32(&db, 0x15, d.di_);
      put_f32(&db, 0x1d, d.dj_);
      body.push_back((char)0x1a);
      put_varint(&body, db.size());
      body += db;
    }
    out.push_back((char)0x0a);
    put_varint(&out, body.size());
    out += body;
  }
  output->write(out.data(), (std::streamsize)out.size());
  return output->good();
}
