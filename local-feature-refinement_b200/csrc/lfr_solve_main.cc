// lfr_solve_main.cc — native `solve` executable: the command-line surface of the reference's
// multi-view-refinement/solve.cc main() (solve.cc:375-682) above the C ABI.
//
//     solve --matches_file X.pb --output_file Y.pb [--n_threads 8] [--banned_images NAME]...
//
// Same flags (solve.cc:379-385, Boost.Program_options semantics and messages), same `.part.N` handling
// (:416-424), same stdout lines (:484-485,534,549,589,591,606,638,641,670) and exit codes (0; 1 on a
// command-line error, :397-401; -1 = 255 when the input does not parse or the output cannot be
// written, :433-436,674-677).  The work is three library calls: lfr_wire_decode_matches()
// (include/lfr_wire.h), lfr_host_stage_create() (include/lfr_host.h) and lfr_solve() /
// lfr_solve_multi() (include/lfr.h) on page-locked arrays, then lfr_wire_encode_solution().
// Opt-in extras: --device D, --gpus N (one process drives N GPUs).  `--n_threads` is accepted and
// ignored (the solve runs on the GPU).  local-feature-refinement_b200/cli.py is the same program in
// Python; this one starts in milliseconds instead of importing numpy.
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fcntl.h>
#include <fstream>
#include <map>
#include <set>
#include <string>
#include <thread>
#include <vector>

#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "../../include/lfr.h"
#include "../../include/lfr_host.h"
#include "../../include/lfr_wire.h"

namespace {

const char* kUsage =
    "Options:\n"
    "  --help                    print the help\n"
    "  --matches_file arg        path to the matches file\n"
    "  --output_file arg         path to the output file\n"
    "  --n_threads arg (=8)      # threads\n"
    "  --banned_images arg (={}) banned images\n";

struct CliError {
  std::string msg;
};

struct OptionSpec {
  const char* name;
  bool takes, repeat;
};
// solve.cc:379-385 plus the opt-in extras
const OptionSpec kOptions[] = {{"help", false, false},        {"matches_file", true, false}, {"output_file", true, false},
                               {"n_threads", true, false},    {"banned_images", true, true}, {"device", true, false},
                               {"gpus", true, false}};

// exact name, else unambiguous prefix (Boost.Program_options' default allow_guessing)
const OptionSpec& find_option(const std::string& name) {
  const OptionSpec* hit = nullptr;
  int n_hits = 0;
  for (const OptionSpec& o : kOptions) {
    if (name == o.name) return o;
    if (std::strncmp(o.name, name.c_str(), name.size()) == 0) {
      hit = &o;
      ++n_hits;
    }
  }
  if (n_hits == 1) return *hit;
  if (n_hits > 1) throw CliError{"option '--" + name + "' is ambiguous"};
  throw CliError{"unrecognised option '--" + name + "'"};
}

struct Args {
  bool help = false;
  std::map<std::string, std::string> single;
  std::vector<std::string> banned;
};

// Boost.Program_options semantics of solve.cc:387-401: long options, `--name value` or `--name=value`,
// no positional arguments, Boost's error messages
Args parse(int argc, char** argv) {
  Args a;
  for (int i = 1; i < argc;) {
    const std::string tok = argv[i++];
    if (tok.size() > 2 && tok.compare(0, 2, "--") == 0) {
      const size_t eq = tok.find('=');
      const std::string name = tok.substr(2, eq == std::string::npos ? std::string::npos : eq - 2);
      const OptionSpec& o = find_option(name);
      if (o.takes) {
        std::string v;
        if (eq != std::string::npos) {
          v = tok.substr(eq + 1);
        } else if (i < argc && !(argv[i][0] == '-' && argv[i][1] != '\0')) {
          v = argv[i++];
        } else {
          throw CliError{std::string("the required argument for option '--") + o.name + "' is missing"};
        }
        if (o.repeat) {
          a.banned.push_back(v);
        } else if (a.single.count(o.name)) {
          throw CliError{std::string("option '--") + o.name + "' cannot be specified more than once"};
        } else {
          a.single[o.name] = v;
        }
      } else {
        if (eq != std::string::npos) throw CliError{std::string("option '--") + o.name + "' does not take any arguments"};
        a.help = true;
      }
    } else if (tok.size() > 1 && tok[0] == '-') {
      throw CliError{"unrecognised option '" + tok + "'"};
    } else {
      throw CliError{"too many positional options have been specified on the command line"};
    }
  }
  return a;
}

uint64_t as_uint(const Args& a, const char* name, uint64_t dflt) {  // lexical_cast<size_t>
  const auto it = a.single.find(name);
  if (it == a.single.end()) return dflt;
  const std::string& v = it->second;
  bool digits = !v.empty();
  for (char c : v) digits = digits && c >= '0' && c <= '9';
  if (!digits) throw CliError{"the argument ('" + v + "') for option '--" + name + "' is invalid"};
  return std::strtoull(v.c_str(), nullptr, 10);
}

bool file_exists(const std::string& path) {
  std::ifstream f(path, std::ios::binary);
  return f.good();
}

// A file mapped read-only (a Madrid-scale MatchingFile is several GB: no second copy of it in memory).
struct FileImage {
  const uint8_t* p = nullptr;
  size_t n = 0;
  bool open(const std::string& path) {
    const int fd = ::open(path.c_str(), O_RDONLY);
    if (fd < 0) return false;
    struct stat st;
    if (fstat(fd, &st) != 0) {
      ::close(fd);
      return false;
    }
    n = (size_t)st.st_size;
    if (n) {
      void* m = mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0);
      ::close(fd);
      if (m == MAP_FAILED) {
        n = 0;
        return false;
      }
      p = static_cast<const uint8_t*>(m);
    } else {
      ::close(fd);
    }
    return true;
  }
  void release() {
    if (p) munmap(const_cast<uint8_t*>(p), n);
    p = nullptr;
    n = 0;
  }
  const uint8_t* data() const { return p; }
  size_t size() const { return n; }
};

// One decoded MatchingFile part, image names still as byte ranges of its buffer.
struct Part {
  FileImage buf;
  uint64_t P = 0, M = 0;
  std::vector<uint64_t> pair_ptr, n1_off, n2_off;
  std::vector<uint32_t> n1_len, n2_len, feat1, feat2;
  std::vector<float> fact1, fact2, sim, disp1, disp2;
};

using Clock = std::chrono::steady_clock;
long long ms_between(Clock::time_point a, Clock::time_point b) {
  return (long long)std::chrono::duration_cast<std::chrono::milliseconds>(b - a).count();
}

}  // namespace

int main(int argc, char** argv) {
  std::string matches_file, output_file;
  std::vector<std::string> banned_list;
  uint64_t device = 0, gpus = 1;
  try {
    const Args a = parse(argc, argv);
    if (a.help) {
      std::fputs("Patch Match graph problem solver\n\n", stdout);
      std::fputs(kUsage, stdout);
      return 0;
    }
    for (const char* req : {"matches_file", "output_file"})  // po::notify: required options, in declaration order
      if (!a.single.count(req)) throw CliError{std::string("the option '--") + req + "' is required but missing"};
    matches_file = a.single.at("matches_file");
    output_file = a.single.at("output_file");
    (void)as_uint(a, "n_threads", 8);
    device = as_uint(a, "device", 0);
    gpus = as_uint(a, "gpus", 1);
    banned_list = a.banned;
  } catch (const CliError& e) {  // solve.cc:397-401
    std::fprintf(stderr, "ERROR: %s\n\n", e.msg.c_str());
    std::fputs(kUsage, stderr);
    return 1;
  }

  // A one-shot process pays for CUDA's start-up (driver initialisation over every visible GPU, context
  // creation: 0.4-0.8 s on an 8-GPU box) more than for all of its work on a Fountain-scale scene.  Two
  // measures: only the device that will be used is made visible (unless the caller already chose),
  // and the start-up runs on a helper thread while the input is read, decoded and turned into a graph.
  if (gpus <= 1 && !std::getenv("CUDA_VISIBLE_DEVICES")) {
    setenv("CUDA_VISIBLE_DEVICES", std::to_string(device).c_str(), 1);
    device = 0;
  }
  std::thread cuda_start([] { lfr_host_free(lfr_host_alloc(64)); });
  struct Joiner {
    std::thread& t;
    ~Joiner() { if (t.joinable()) t.join(); }
  } join_on_exit{cuda_start};

  // ---- solve.cc:412-481: the file, or its .part.N pieces ------------------------------------
  std::vector<std::string> files;
  if (file_exists(matches_file)) {
    files.push_back(matches_file);
  } else {
    for (int k = 0;; ++k) {
      const std::string f = matches_file + ".part." + std::to_string(k);
      if (!file_exists(f)) break;
      files.push_back(f);
    }
  }
  std::vector<Part> parts(files.size());
  for (size_t i = 0; i < files.size(); ++i) {
    Part& pt = parts[i];
    bool ok = pt.buf.open(files[i]);
    if (ok) ok = lfr_wire_scan_matches(pt.buf.data(), pt.buf.size(), &pt.P, &pt.M) == LFR_OK;
    if (ok) {
      pt.pair_ptr.resize(pt.P + 1);
      pt.n1_off.resize(pt.P); pt.n2_off.resize(pt.P); pt.n1_len.resize(pt.P); pt.n2_len.resize(pt.P);
      pt.fact1.resize(pt.P); pt.fact2.resize(pt.P);
      pt.feat1.resize(pt.M); pt.feat2.resize(pt.M); pt.sim.resize(pt.M);
      pt.disp1.resize(18 * pt.M); pt.disp2.resize(18 * pt.M);
      lfr_wire_matches w;
      w.n_pairs = pt.P; w.n_matches = pt.M;
      w.pair_ptr = pt.pair_ptr.data(); w.fact1 = pt.fact1.data(); w.fact2 = pt.fact2.data();
      w.name1_off = pt.n1_off.data(); w.name1_len = pt.n1_len.data();
      w.name2_off = pt.n2_off.data(); w.name2_len = pt.n2_len.data();
      w.feat1 = pt.feat1.data(); w.feat2 = pt.feat2.data(); w.sim = pt.sim.data();
      w.disp1 = pt.disp1.data(); w.disp2 = pt.disp2.data();
      ok = lfr_wire_decode_matches(pt.buf.data(), pt.buf.size(), &w) == LFR_OK;
    }
    if (!ok) {
      std::fputs("Failed to parse proto object.\n", stderr);  // solve.cc:433-436
      return 255;
    }
  }

  // image names -> ids in order of first appearance over all parts; the pairs of all parts in order
  std::vector<std::string> names;
  std::map<std::string, uint32_t> id_of;
  auto intern = [&](const Part& pt, uint64_t off, uint32_t len) {
    std::string s(reinterpret_cast<const char*>(pt.buf.data()) + off, len);
    const auto it = id_of.find(s);
    if (it != id_of.end()) return it->second;
    const uint32_t id = (uint32_t)names.size();
    id_of.emplace(s, id);
    names.push_back(std::move(s));
    return id;
  };
  uint64_t P = 0, M = 0;
  for (const Part& pt : parts) { P += pt.P; M += pt.M; }
  std::vector<uint32_t> pair_img1(P), pair_img2(P);
  std::vector<float> pair_fact1(P), pair_fact2(P);
  std::vector<uint64_t> pair_ptr(P + 1, 0);
  // single part: the decoded arrays are used in place; several: concatenated
  std::vector<uint32_t> feat1_cat, feat2_cat;
  std::vector<float> sim_cat, disp1_cat, disp2_cat;
  const bool cat = parts.size() > 1;
  if (cat) {
    feat1_cat.reserve(M); feat2_cat.reserve(M); sim_cat.reserve(M);
    disp1_cat.reserve(18 * M); disp2_cat.reserve(18 * M);
  }
  {
    uint64_t p0 = 0, m0 = 0;
    for (Part& pt : parts) {
      for (uint64_t p = 0; p < pt.P; ++p) {
        pair_img1[p0 + p] = intern(pt, pt.n1_off[p], pt.n1_len[p]);
        pair_img2[p0 + p] = intern(pt, pt.n2_off[p], pt.n2_len[p]);
        pair_fact1[p0 + p] = pt.fact1[p];
        pair_fact2[p0 + p] = pt.fact2[p];
        pair_ptr[p0 + p + 1] = m0 + pt.pair_ptr[p + 1];
      }
      if (cat) {
        feat1_cat.insert(feat1_cat.end(), pt.feat1.begin(), pt.feat1.end());
        feat2_cat.insert(feat2_cat.end(), pt.feat2.begin(), pt.feat2.end());
        sim_cat.insert(sim_cat.end(), pt.sim.begin(), pt.sim.end());
        disp1_cat.insert(disp1_cat.end(), pt.disp1.begin(), pt.disp1.end());
        disp2_cat.insert(disp2_cat.end(), pt.disp2.begin(), pt.disp2.end());
      }
      p0 += pt.P;
      m0 += pt.M;
      pt.buf.release();  // the file image (several GB at Madrid scale) is not needed past the names
      if (cat) {
        std::vector<uint32_t>().swap(pt.feat1);
        std::vector<uint32_t>().swap(pt.feat2);
        std::vector<float>().swap(pt.sim);
        std::vector<float>().swap(pt.disp1);
        std::vector<float>().swap(pt.disp2);
      }
    }
  }
  const uint32_t* feat1 = cat ? feat1_cat.data() : (parts.empty() ? nullptr : parts[0].feat1.data());
  const uint32_t* feat2 = cat ? feat2_cat.data() : (parts.empty() ? nullptr : parts[0].feat2.data());
  const float* sim = cat ? sim_cat.data() : (parts.empty() ? nullptr : parts[0].sim.data());
  const float* disp1 = cat ? disp1_cat.data() : (parts.empty() ? nullptr : parts[0].disp1.data());
  const float* disp2 = cat ? disp2_cat.data() : (parts.empty() ? nullptr : parts[0].disp2.data());

  // banned images (solve.cc:444-446) and the first fact seen of every image among the pairs kept
  const std::set<std::string> banned(banned_list.begin(), banned_list.end());
  std::vector<uint8_t> skip(P, 0);
  std::vector<float> image_fact(names.size(), 0.f);
  std::vector<uint8_t> have_fact(names.size(), 0);
  uint64_t e_cap = 0;
  for (uint64_t p = 0; p < P; ++p) {
    skip[p] = (banned.count(names[pair_img1[p]]) || banned.count(names[pair_img2[p]])) ? 1 : 0;
    if (skip[p]) continue;
    e_cap += 2 * (pair_ptr[p + 1] - pair_ptr[p]);
    if (!have_fact[pair_img1[p]]) { have_fact[pair_img1[p]] = 1; image_fact[pair_img1[p]] = pair_fact1[p]; }
    if (!have_fact[pair_img2[p]]) { have_fact[pair_img2[p]] = 1; image_fact[pair_img2[p]] = pair_fact2[p]; }
  }

  cuda_start.join();  // page-locked memory needs the context
  const Clock::time_point t_start = Clock::now();  // "Total time" (solve.cc:487-641): the graph work and the solve
  // ---- host graph stage (solve.cc:438-606); the edge records go straight into page-locked memory
  lfr_edge* edges = nullptr;
  bool edges_pinned = false;
  if (e_cap) {
    edges = static_cast<lfr_edge*>(lfr_host_alloc(e_cap * sizeof(lfr_edge)));
    edges_pinned = edges != nullptr;
    if (!edges) edges = static_cast<lfr_edge*>(std::malloc(e_cap * sizeof(lfr_edge)));
  }
  lfr_host_input in;
  std::memset(&in, 0, sizeof in);
  in.n_pairs = P; in.n_matches = M; in.n_images = (uint32_t)names.size();
  in.pair_img1 = pair_img1.data(); in.pair_img2 = pair_img2.data(); in.pair_skip = skip.data();
  in.pair_ptr = pair_ptr.data(); in.feat1 = feat1; in.feat2 = feat2; in.sim = sim; in.disp1 = disp1; in.disp2 = disp2;
  in.edges_out = edges; in.edges_out_capacity = e_cap;
  lfr_host_stage* hs = nullptr;
  lfr_host_sizes sz;
  if (lfr_host_stage_create(&in, &hs, &sz) != LFR_OK) {
    std::fprintf(stderr, "ERROR: the host graph stage rejected the input\n");
    return 2;
  }
  const uint32_t N = sz.n_nodes, C = sz.n_components;
  const uint64_t E = sz.n_edges;
  std::vector<uint32_t> row_ptr((size_t)N + 1, 0), track(N), comp(N), comp_ptr((size_t)C + 1, 0), comp_nodes(N), comp_order(C),
      node_image(N), node_feat(N);
  std::vector<uint8_t> is_root(N);
  if (lfr_host_stage_export(hs, row_ptr.data(), edges, track.data(), comp.data(), is_root.data(), comp_ptr.data(),
                            comp_nodes.data(), comp_order.data(), node_image.data(), node_feat.data()) != LFR_OK) {
    std::fprintf(stderr, "ERROR: lfr_host_stage_export failed\n");
    return 2;
  }
  lfr_host_stage_destroy(hs);
  std::printf("# graph nodes: %u\n", N);                                  // solve.cc:484
  std::printf("# graph edges: %llu\n", (unsigned long long)E);            // solve.cc:485
  if (N) {
    std::printf("# tracks: %u\n", sz.n_tracks);                           // solve.cc:534
    std::printf("max track size: %u\n", sz.max_track_size);               // solve.cc:549
    std::printf("Graph-cut time: %lldms\n", (long long)sz.graph_cut_ms);  // solve.cc:589
    std::printf("# components: %u\n", C);                                 // solve.cc:591
    std::printf("max component size: %u\n", sz.max_component_size);       // solve.cc:606
  }
  std::fflush(stdout);

  // ---- the solve (solve.cc:608-635): positions start at zero --------------------------------
  double* positions = nullptr;
  bool positions_pinned = false;
  if (N) {
    positions = static_cast<double*>(lfr_host_alloc(2ull * N * sizeof(double)));
    positions_pinned = positions != nullptr;
    if (!positions) positions = static_cast<double*>(std::malloc(2ull * N * sizeof(double)));
    std::memset(positions, 0, 2ull * N * sizeof(double));
  }
  const Clock::time_point t_solve = Clock::now();
  if (N) {
    lfr_problem prob;
    prob.n_nodes = N; prob.n_components = C; prob.n_edges = E;
    prob.row_ptr = row_ptr.data(); prob.edges = edges; prob.track = track.data(); prob.comp = comp.data();
    prob.is_root = is_root.data(); prob.comp_ptr = comp_ptr.data(); prob.comp_nodes = comp_nodes.data();
    lfr_options opt;
    lfr_options_default(&opt);
    opt.device = (int32_t)device;
    lfr_stats st;
    std::memset(&st, 0, sizeof st);
    int rc;
    if (gpus > 1) {
      std::vector<int32_t> devs((size_t)gpus);
      for (size_t i = 0; i < devs.size(); ++i) devs[i] = (int32_t)i;
      rc = lfr_solve_multi(&prob, &opt, devs.data(), (int32_t)devs.size(), positions, &st, nullptr);
    } else {
      rc = lfr_solve(&prob, &opt, positions, &st);
    }
    if (rc != LFR_OK) {  // no CPU fallback: without a CUDA device this is where the program stops
      std::fprintf(stderr, "ERROR: the solve failed (%d): %s\n", rc, lfr_last_error());
      return 2;
    }
  }
  const Clock::time_point t_end = Clock::now();
  std::printf("Solver time: %lldms\n", ms_between(t_solve, t_end));  // solve.cc:638
  std::printf("Total time: %lldms\n", ms_between(t_start, t_end));   // solve.cc:641

  // ---- SolutionFile (solve.cc:643-679): images by first node appearance, nodes by index ------
  std::vector<int64_t> rank(names.size(), -1);
  std::vector<uint32_t> img_ids;
  for (uint32_t v = 0; v < N; ++v)
    if (rank[node_image[v]] < 0) {
      rank[node_image[v]] = (int64_t)img_ids.size();
      img_ids.push_back(node_image[v]);
    }
  const size_t I = img_ids.size();
  std::vector<uint64_t> img_ptr(I + 1, 0);
  for (uint32_t v = 0; v < N; ++v) ++img_ptr[(size_t)rank[node_image[v]] + 1];
  for (size_t i = 0; i < I; ++i) img_ptr[i + 1] += img_ptr[i];
  std::vector<uint32_t> feature_idx(N);
  std::vector<float> di(N), dj(N);
  uint64_t n_outside = 0;
  {
    std::vector<uint64_t> fill(img_ptr.begin(), img_ptr.end() - 1);
    for (uint32_t v = 0; v < N; ++v) {
      const uint64_t k = fill[(size_t)rank[node_image[v]]]++;
      feature_idx[k] = node_feat[v];
      di[k] = (float)positions[2 * (size_t)v];
      dj[k] = (float)positions[2 * (size_t)v + 1];
      if (std::fabs(positions[2 * (size_t)v + 1]) > 0.5 || std::fabs(positions[2 * (size_t)v]) > 0.5) ++n_outside;  // solve.cc:666-669
    }
  }
  std::printf("# points with at least one coordinate > 0.5: %llu\n", (unsigned long long)n_outside);  // solve.cc:670
  std::fflush(stdout);
  std::vector<uint8_t> name_blob;
  std::vector<uint64_t> name_off(I + 1, 0);
  std::vector<float> fact(I);
  for (size_t i = 0; i < I; ++i) {
    const std::string& s = names[img_ids[i]];
    name_blob.insert(name_blob.end(), s.begin(), s.end());
    name_off[i + 1] = name_blob.size();
    fact[i] = image_fact[img_ids[i]];
  }
  if (name_blob.empty()) name_blob.push_back(0);
  const int64_t need = lfr_wire_encode_solution(I, img_ptr.data(), name_blob.data(), name_off.data(), fact.data(),
                                                feature_idx.data(), di.data(), dj.data(), nullptr, 0);
  std::vector<uint8_t> out((size_t)std::max<int64_t>(need, 1));
  bool written = need >= 0 && lfr_wire_encode_solution(I, img_ptr.data(), name_blob.data(), name_off.data(), fact.data(),
                                                       feature_idx.data(), di.data(), dj.data(), out.data(), (uint64_t)need) == need;
  if (written) {
    std::ofstream f(output_file, std::ios::binary | std::ios::trunc);
    written = f.good();
    if (written && need) f.write(reinterpret_cast<const char*>(out.data()), need);
    written = written && f.good();
  }
  if (positions_pinned) lfr_host_free(positions); else std::free(positions);
  if (edges_pinned) lfr_host_free(edges); else std::free(edges);
  if (!written) {
    std::fputs("Failed to write proto object.\n", stderr);  // solve.cc:674-677
    return 255;
  }
  return 0;
}
