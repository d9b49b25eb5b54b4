// Stand-in for cxxopts.hpp  --  TEST INFRASTRUCTURE ONLY (see Eigen/Dense in this directory).
//
// The subset of cxxopts' public interface used by /root/reference/c++/src/simpleicp-cli.cpp:
// Options(name, help), add_options()("s,long", "description", value<T>()->default_value("..")),
// parse(argc, argv) -> result.count("long"), result["long"].as<T>(), help().
#ifndef SICP_ORACLE_CXXOPTS_STANDIN
#define SICP_ORACLE_CXXOPTS_STANDIN

#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace cxxopts
{

class Value : public std::enable_shared_from_this<Value>
{
public:
  explicit Value(bool flag) : flag_(flag), has_default_(false) {}
  std::shared_ptr<Value> default_value(const std::string &v)
  {
    has_default_ = true;
    default_ = v;
    return shared_from_this();
  }
  bool flag_;
  bool has_default_;
  std::string default_;
};

template <typename T> std::shared_ptr<Value> value() { return std::make_shared<Value>(false); }

class OptionValue
{
public:
  OptionValue() : present_(false), count_(0) {}
  template <typename T> T as() const
  {
    if (!present_)
      throw std::runtime_error("Option '" + name_ + "' has no value");
    T out;
    convert(text_, out);
    return out;
  }
  std::string name_, text_;
  bool present_;
  size_t count_;

private:
  static void convert(const std::string &s, std::string &o) { o = s; }
  template <typename T> static void convert(const std::string &s, T &o)
  {
    std::istringstream in(s);
    in >> o;
    if (in.fail() || !in.eof())
      throw std::runtime_error("Argument '" + s + "' failed to parse");
  }
};

class ParseResult
{
public:
  size_t count(const std::string &name) const
  {
    auto it = values_.find(name);
    return it == values_.end() ? 0 : it->second.count_;
  }
  const OptionValue &operator[](const std::string &name) const
  {
    auto it = values_.find(name);
    if (it == values_.end())
      throw std::runtime_error("Option '" + name + "' does not exist");
    return it->second;
  }
  std::map<std::string, OptionValue> values_;
};

class Options;

class OptionAdder
{
public:
  explicit OptionAdder(Options &o) : o_(o) {}
  OptionAdder &operator()(const std::string &opts, const std::string &desc,
                          const std::shared_ptr<Value> &value = std::make_shared<Value>(true));

private:
  Options &o_;
};

class Options
{
public:
  Options(std::string program, std::string help_string) : program_(std::move(program)), help_(std::move(help_string)) {}
  OptionAdder add_options() { return OptionAdder(*this); }

  ParseResult parse(int argc, char **argv)
  {
    ParseResult r;
    for (auto &d : defs_)
    {
      OptionValue v;
      v.name_ = d.long_name;
      if (d.value->has_default_)
      {
        v.present_ = true;
        v.text_ = d.value->default_;
      }
      r.values_[d.long_name] = v;
    }
    for (int i = 1; i < argc; i++)
    {
      std::string a = argv[i], name, text;
      bool has_text = false;
      if (a.rfind("--", 0) == 0)
      {
        name = a.substr(2);
        auto eq = name.find('=');
        if (eq != std::string::npos)
        {
          text = name.substr(eq + 1);
          name = name.substr(0, eq);
          has_text = true;
        }
      }
      else if (a.size() == 2 && a[0] == '-')
      {
        for (auto &d : defs_)
          if (d.short_name == a.substr(1))
            name = d.long_name;
        if (name.empty())
          throw std::runtime_error("Option '" + a.substr(1) + "' does not exist");
      }
      else
        throw std::runtime_error("Unexpected argument '" + a + "'");
      const Def *def = nullptr;
      for (auto &d : defs_)
        if (d.long_name == name)
          def = &d;
      if (!def)
        throw std::runtime_error("Option '" + name + "' does not exist");
      OptionValue &v = r.values_[name];
      v.count_++;
      if (def->value->flag_)
      {
        v.present_ = true;
        v.text_ = "true";
        continue;
      }
      if (!has_text)
      {
        if (i + 1 >= argc)
          throw std::runtime_error("Option '" + name + "' is missing an argument");
        text = argv[++i];
      }
      v.present_ = true;
      v.text_ = text;
    }
    return r;
  }

  std::string help() const
  {
    std::ostringstream out;
    out << help_ << "\nUsage:\n  " << program_ << " [OPTION...]\n\n";
    for (auto &d : defs_)
    {
      out << "  ";
      if (!d.short_name.empty())
        out << "-" << d.short_name << ", ";
      out << "--" << d.long_name;
      if (!d.value->flag_)
        out << " arg";
      out << "  " << d.desc;
      if (d.value->has_default_)
        out << " (default: " << d.value->default_ << ")";
      out << "\n";
    }
    return out.str();
  }

  struct Def
  {
    std::string short_name, long_name, desc;
    std::shared_ptr<Value> value;
  };
  std::vector<Def> defs_;

private:
  std::string program_, help_;
};

inline OptionAdder &OptionAdder::operator()(const std::string &opts, const std::string &desc,
                                            const std::shared_ptr<Value> &value)
{
  Options::Def d;
  auto comma = opts.find(',');
  if (comma == std::string::npos)
    d.long_name = opts;
  else
  {
    d.short_name = opts.substr(0, comma);
    d.long_name = opts.substr(comma + 1);
  }
  d.desc = desc;
  d.value = value;
  o_.defs_.push_back(d);
  return *this;
}

} // namespace cxxopts

#endif // SICP_ORACLE_CXXOPTS_STANDIN
