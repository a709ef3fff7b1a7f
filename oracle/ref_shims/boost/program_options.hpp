// boost/program_options.hpp — SHIM (test infrastructure written for this repo; not Boost).
//
// The slice of Boost.Program_options that multi-view-refinement/solve.cc:379-403,405,413,617
// uses, with Boost's command-line semantics for it: long options `--name value` / `--name=value`,
// unambiguous prefixes accepted (allow_guessing), typed values, required(), default_value(v, text),
// vector<string> options composing over repeated occurrences, `po::error` exceptions with Boost's
// messages, and the help table printed by `std::cout << desc`.
#ifndef LFR_SHIM_BOOST_PROGRAM_OPTIONS_
#define LFR_SHIM_BOOST_PROGRAM_OPTIONS_

#include <cstddef>
#include <cstdlib>
#include <map>
#include <memory>
#include <ostream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <typeinfo>
#include <vector>

namespace boost {
namespace program_options {

class error : public std::logic_error {
 public:
  explicit error(const std::string& what) : std::logic_error(what) {}
};

// a parsed value of any type
class variable_value {
 public:
  variable_value() : defaulted_(false) {}
  bool empty() const { return !holder_; }
  bool defaulted() const { return defaulted_; }
  template <typename T>
  const T& as() const {
    if (!holder_ || *type_ != typeid(T)) throw error("boost::bad_any_cast: failed conversion using boost::any_cast");
    return *static_cast<const T*>(holder_.get());
  }
  template <typename T>
  T& mutable_as() {
    return *static_cast<T*>(holder_.get());
  }
  template <typename T>
  void set(const T& v, bool defaulted) {
    holder_ = std::shared_ptr<void>(new T(v), [](void* p) { delete static_cast<T*>(p); });
    type_ = &typeid(T);
    defaulted_ = defaulted;
  }

 private:
  std::shared_ptr<void> holder_;
  const std::type_info* type_ = nullptr;
  bool defaulted_;
};

class value_semantic {
 public:
  virtual ~value_semantic() {}
  virtual bool is_required() const = 0;
  virtual bool takes_argument() const = 0;
  virtual std::string name() const = 0;                                            // "arg", "arg (=8)"
  virtual void apply_default(variable_value& v) const = 0;                         // no-op without a default
  virtual void parse(variable_value& v, const std::string& token, const std::string& option) const = 0;
};

namespace detail {
template <typename T>
struct lexical {
  static T cast(const std::string& s, const std::string& option) {
    std::istringstream in(s);
    T v;
    in >> v;
    if (in.fail() || !in.eof()) throw error("the argument ('" + s + "') for option '--" + option + "' is invalid");
    return v;
  }
};
template <>
struct lexical<std::string> {
  static std::string cast(const std::string& s, const std::string&) { return s; }
};
template <typename T>
struct is_vector {
  enum { value = 0 };
};
template <typename U>
struct is_vector<std::vector<U> > {
  enum { value = 1 };
};
}  // namespace detail

template <typename T>
class typed_value : public value_semantic {
 public:
  typed_value() : required_(false), has_default_(false) {}
  typed_value* required() {
    required_ = true;
    return this;
  }
  typed_value* default_value(const T& v, const std::string& textual) {
    default_ = v;
    default_text_ = textual;
    has_default_ = true;
    return this;
  }
  virtual bool is_required() const { return required_; }
  virtual bool takes_argument() const { return true; }
  virtual std::string name() const { return has_default_ ? "arg (=" + default_text_ + ")" : "arg"; }
  virtual void apply_default(variable_value& v) const {
    if (has_default_ && v.empty()) v.set<T>(default_, true);
  }
  virtual void parse(variable_value& v, const std::string& token, const std::string& option) const {
    parse_impl(v, token, option, static_cast<T*>(nullptr));
  }

 private:
  template <typename U>
  void parse_impl(variable_value& v, const std::string& token, const std::string& option, U*) const {
    if (!v.empty() && !v.defaulted()) throw error("option '--" + option + "' cannot be specified more than once");
    v.set<U>(detail::lexical<U>::cast(token, option), false);
  }
  template <typename U>
  void parse_impl(variable_value& v, const std::string& token, const std::string& option, std::vector<U>*) const {
    // composing: every occurrence appends one element
    if (v.empty() || v.defaulted()) v.set<std::vector<U> >(std::vector<U>(), false);
    v.mutable_as<std::vector<U> >().push_back(detail::lexical<U>::cast(token, option));
  }
  bool required_, has_default_;
  T default_;
  std::string default_text_;
};

template <typename T>
typed_value<T>* value() {
  return new typed_value<T>();
}

struct option_description {
  std::string name, description;
  std::shared_ptr<const value_semantic> semantic;  // null: a switch without argument
};

class options_description;

class options_description_easy_init {
 public:
  explicit options_description_easy_init(options_description* owner) : owner_(owner) {}
  options_description_easy_init& operator()(const char* name, const char* description);
  options_description_easy_init& operator()(const char* name, const value_semantic* s, const char* description);

 private:
  options_description* owner_;
};

class options_description {
 public:
  explicit options_description(const std::string& caption) : caption_(caption) {}
  options_description_easy_init add_options() { return options_description_easy_init(this); }
  void add(const option_description& d) { options_.push_back(d); }
  const std::vector<option_description>& options() const { return options_; }
  const std::string& caption() const { return caption_; }
  // exact name, else unambiguous prefix (Boost's allow_guessing)
  const option_description& find(const std::string& name) const {
    const option_description* hit = nullptr;
    int n_prefix = 0;
    for (size_t i = 0; i < options_.size(); ++i) {
      if (options_[i].name == name) return options_[i];
      if (options_[i].name.compare(0, name.size(), name) == 0) {
        hit = &options_[i];
        ++n_prefix;
      }
    }
    if (n_prefix == 1) return *hit;
    if (n_prefix > 1) throw error("option '--" + name + "' is ambiguous");
    throw error("unrecognised option '--" + name + "'");
  }

 private:
  std::string caption_;
  std::vector<option_description> options_;
};

inline options_description_easy_init& options_description_easy_init::operator()(const char* name, const char* description) {
  option_description d;
  d.name = name;
  d.description = description;
  owner_->add(d);
  return *this;
}
inline options_description_easy_init& options_description_easy_init::operator()(const char* name, const value_semantic* s,
                                                                                 const char* description) {
  option_description d;
  d.name = name;
  d.description = description;
  d.semantic.reset(s);
  owner_->add(d);
  return *this;
}

inline std::ostream& operator<<(std::ostream& os, const options_description& desc) {
  // Boost's layout: caption, then "  --name arg" padded to the description column
  // (widest first column + 1, at least 23... ) — Boost computes the column as the longest
  // "  --name arg" + 1, capped by the line length (80).
  std::vector<std::string> first(desc.options().size());
  size_t width = 23;
  for (size_t i = 0; i < desc.options().size(); ++i) {
    const option_description& o = desc.options()[i];
    first[i] = "  --" + o.name + (o.semantic ? " " + o.semantic->name() : "");
    width = std::max(width, first[i].size());
  }
  ++width;
  os << desc.caption() << ":\n";
  for (size_t i = 0; i < desc.options().size(); ++i) {
    os << first[i];
    for (size_t pad = first[i].size(); pad < width; ++pad) os.put(' ');
    os << desc.options()[i].description << "\n";
  }
  return os;
}

struct parsed_options {
  const options_description* description;
  std::vector<std::pair<std::string, std::vector<std::string> > > options;  // canonical name -> tokens ("" for switches)
};

inline parsed_options parse_command_line(int argc, const char* const* argv, const options_description& desc) {
  parsed_options out;
  out.description = &desc;
  for (int i = 1; i < argc; ++i) {
    const std::string tok = argv[i];
    if (tok.size() >= 2 && tok[0] == '-' && tok[1] == '-') {
      std::string name = tok.substr(2), adjacent;
      bool has_adjacent = false;
      const size_t eq = name.find('=');
      if (eq != std::string::npos) {
        adjacent = name.substr(eq + 1);
        name = name.substr(0, eq);
        has_adjacent = true;
      }
      const option_description& d = desc.find(name);
      std::vector<std::string> values;
      if (d.semantic && d.semantic->takes_argument()) {
        if (has_adjacent) {
          values.push_back(adjacent);
        } else if (i + 1 < argc && !(argv[i + 1][0] == '-' && argv[i + 1][1] != '\0')) {
          values.push_back(argv[++i]);
        } else {
          throw error("the required argument for option '--" + d.name + "' is missing");
        }
      } else if (has_adjacent) {
        throw error("option '--" + d.name + "' does not take any arguments");
      }
      out.options.push_back(std::make_pair(d.name, values));
    } else if (tok.size() >= 2 && tok[0] == '-') {
      throw error("unrecognised option '" + tok + "'");
    } else {
      throw error("too many positional options have been specified on the command line");
    }
  }
  return out;
}

class variables_map {
 public:
  size_t count(const std::string& name) const {
    std::map<std::string, variable_value>::const_iterator it = values_.find(name);
    return (it != values_.end() && !it->second.empty()) ? 1 : 0;
  }
  const variable_value& operator[](const std::string& name) const {
    static const variable_value empty;
    std::map<std::string, variable_value>::const_iterator it = values_.find(name);
    return it == values_.end() ? empty : it->second;
  }
  std::map<std::string, variable_value> values_;
  const options_description* description_ = nullptr;
};

inline void store(const parsed_options& parsed, variables_map& vm) {
  vm.description_ = parsed.description;
  for (size_t i = 0; i < parsed.options.size(); ++i) {
    const option_description& d = parsed.description->find(parsed.options[i].first);
    variable_value& v = vm.values_[d.name];
    if (!d.semantic) {
      v.set<bool>(true, false);
      continue;
    }
    for (size_t k = 0; k < parsed.options[i].second.size(); ++k) d.semantic->parse(v, parsed.options[i].second[k], d.name);
  }
  for (size_t i = 0; i < parsed.description->options().size(); ++i) {  // defaults for what was not given
    const option_description& d = parsed.description->options()[i];
    if (d.semantic) d.semantic->apply_default(vm.values_[d.name]);
  }
}

inline void notify(variables_map& vm) {
  if (!vm.description_) return;
  for (size_t i = 0; i < vm.description_->options().size(); ++i) {
    const option_description& d = vm.description_->options()[i];
    if (d.semantic && d.semantic->is_required() && !vm.count(d.name))
      throw error("the option '--" + d.name + "' is required but missing");
  }
}

}  // namespace program_options
}  // namespace boost

#endif  // LFR_SHIM_BOOST_PROGRAM_OPTIONS_
