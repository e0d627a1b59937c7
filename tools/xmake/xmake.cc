// xmake -- the repo's own build tool (C++20, no dependencies), the counterpart of the reference's
// build/xmake.cc. Same target files: an `xmake.yml` next to the sources maps
//     target -> {main, srcs, hdrs, deps, rule, lopts, gopts}
// (reference build/xmake.cc:92-103; `rule`, `lopts`, `gopts` exist there but are unused, 280-285), with
// the rule the reference left open filled in:
//
//   rule: c++   (default)  $CXX -O3 -std=c++20 -Wall   -- the reference's flags (xmake.cc:191-205) minus
//                          -mavx -ffast-math: the host only orchestrates, the arithmetic is on the device
//   rule: cuda             $NVCC -gencode arch=compute_100a,code=sm_100a -lineinfo -O3; with
//                          lopts: ["-shared"] the objects are linked into lib<target>.so, which C++
//                          targets depend on through its extern "C" header only (include/dfrl.h)
//
// Usage, from anywhere inside the repo (reference: `xmake <target>` inside the package directory):
//     xmake ppo_training                                       a target of the current package
//     xmake //dependence_free_rl_b200/host/apps/bin_packing/ppo_training
//     xmake -n <target>      print the commands without running them      xmake -B <target>   rebuild all
// Labels are //path/from/repo/root/name, `deps` use the same form. Rebuilds are mtime based, outputs go to
// <package dir>/.out/. The repo root is the closest ancestor directory holding `include/dfrl.h`
// (the reference walks up to `.git`, xmake.cc:15-26; snapshots of this repo travel without `.git`).
//
// Bootstrap (the reference's stage 0 is scripts/build/xmake.py:53-61):  tools/xmake/bootstrap.sh
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <filesystem>
#include <fstream>
#include <map>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace fs = std::filesystem;

namespace {

[[noreturn]] void die(const std::string &msg) {
  std::fprintf(stderr, "xmake: %s\n", msg.c_str());
  std::exit(2);
}

// ------------------------------------------------------------------ the xmake.yml subset --------
// Two levels of block mappings; a value is a scalar, an inline list `[a, "b"]` or a block list of
// `- item` lines. Comments start with '#'. That is everything the reference's target files use.
struct target {
  bool main = false;
  std::string rule = "c++";
  std::vector<std::string> srcs, hdrs, deps, lopts, gopts, data;
};

std::string trim(std::string s) {
  const char *ws = " \t\r\n";
  s.erase(0, s.find_first_not_of(ws));
  s.erase(s.find_last_not_of(ws) + 1);
  return s;
}

std::string unquote(std::string s) {
  s = trim(std::move(s));
  if (s.size() >= 2 && (s.front() == '"' || s.front() == '\'') && s.back() == s.front())
    return s.substr(1, s.size() - 2);
  return s;
}

std::string strip_comment(const std::string &line) {
  char quote = 0;
  for (std::size_t i = 0; i < line.size(); ++i) {
    char c = line[i];
    if (quote) {
      if (c == quote)
        quote = 0;
    } else if (c == '"' || c == '\'') {
      quote = c;
    } else if (c == '#' && (i == 0 || line[i - 1] == ' ' || line[i - 1] == '\t')) {
      return line.substr(0, i);
    }
  }
  return line;
}

std::vector<std::string> inline_list(const std::string &v, const std::string &where) {
  if (v.size() < 2 || v.front() != '[' || v.back() != ']')
    die(where + ": expected [a, b, ...]");
  std::vector<std::string> out;
  std::string item;
  char quote = 0;
  for (char c : v.substr(1, v.size() - 2)) {
    if (quote) {
      item += c;
      if (c == quote)
        quote = 0;
    } else if (c == '"' || c == '\'') {
      quote = c;
      item += c;
    } else if (c == ',') {
      out.push_back(unquote(item));
      item.clear();
    } else {
      item += c;
    }
  }
  if (!trim(item).empty())
    out.push_back(unquote(item));
  return out;
}

std::map<std::string, target> parse_targets(const fs::path &file) {
  std::ifstream in(file);
  if (!in)
    die("cannot read " + file.string());
  std::map<std::string, target> targets;
  target *cur = nullptr;
  std::vector<std::string> *list = nullptr;
  std::string raw;
  int lineno = 0;
  while (std::getline(in, raw)) {
    ++lineno;
    const std::string where = file.string() + ":" + std::to_string(lineno);
    std::string line = strip_comment(raw);
    if (trim(line).empty())
      continue;
    const std::size_t indent = line.find_first_not_of(' ');
    line = trim(line);
    if (line[0] == '-') {  // block-list item of the open key
      if (!list)
        die(where + ": list item outside a list");
      list->push_back(unquote(line.substr(1)));
      continue;
    }
    const std::size_t colon = line.find(':');
    if (colon == std::string::npos)
      die(where + ": expected `key: value`");
    const std::string key = trim(line.substr(0, colon)), value = trim(line.substr(colon + 1));
    if (indent == 0) {  // a new target
      if (!value.empty())
        die(where + ": a target is a mapping");
      cur = &targets[key];
      list = nullptr;
      continue;
    }
    if (!cur)
      die(where + ": field outside a target");
    list = nullptr;
    auto field = [&](std::vector<std::string> &dst) {
      if (value.empty())
        list = &dst;  // block list follows
      else
        dst = inline_list(value, where);
    };
    if (key == "main")
      cur->main = (value == "true" || value == "yes" || value == "1");
    else if (key == "rule")
      cur->rule = unquote(value);
    else if (key == "srcs")
      field(cur->srcs);
    else if (key == "hdrs")
      field(cur->hdrs);
    else if (key == "deps")
      field(cur->deps);
    else if (key == "lopts")
      field(cur->lopts);
    else if (key == "gopts")
      field(cur->gopts);
    else if (key == "data")
      field(cur->data);
    else
      die(where + ": unknown field `" + key + "`");
  }
  return targets;
}

// ------------------------------------------------------------------ the build ------------------
struct product {
  std::vector<std::string> objects, libraries;  // what a dependent links
  fs::file_time_type newest = fs::file_time_type::min();
};

struct builder {
  fs::path root;
  bool dry_run = false, force = false;
  std::string cxx = "g++", nvcc = "nvcc";
  std::map<std::string, product> done;
  std::set<std::string> open;  // dependency-cycle detection
  std::map<std::string, std::map<std::string, target>> packages;

  static fs::file_time_type mtime(const fs::path &p) {
    std::error_code ec;
    auto t = fs::last_write_time(p, ec);
    return ec ? fs::file_time_type::min() : t;
  }

  static std::string shell_quote(const std::string &s) {
    if (!s.empty() && s.find_first_of(" \t'\"$&|;<>()*?[]#~`\\") == std::string::npos)
      return s;
    std::string q = "'";
    for (char c : s)
      q += c == '\'' ? std::string("'\\''") : std::string(1, c);
    return q + "'";
  }

  void run(const std::vector<std::string> &argv) {
    std::string cmd;
    for (const auto &a : argv)
      cmd += (cmd.empty() ? "" : " ") + shell_quote(a);
    std::printf("%s\n", cmd.c_str());
    std::fflush(stdout);
    if (dry_run)
      return;
    int rc = std::system(cmd.c_str());
    if (rc != 0)
      die("command failed (" + std::to_string(rc) + "): " + argv[0]);
  }

  const target &find(const std::string &label, std::string &pkg, std::string &name) {
    if (label.rfind("//", 0) != 0)
      die("target labels look like //pkg/name, got " + label);
    const std::size_t slash = label.rfind('/');
    pkg = label.substr(2, slash - 2);
    name = label.substr(slash + 1);
    auto it = packages.find(pkg);
    if (it == packages.end())
      it = packages.emplace(pkg, parse_targets(root / pkg / "xmake.yml")).first;
    auto t = it->second.find(name);
    if (t == it->second.end())
      die("no target " + name + " in " + pkg + "/xmake.yml");
    return t->second;
  }

  const product &build(const std::string &label) {
    if (auto it = done.find(label); it != done.end())
      return it->second;
    if (!open.insert(label).second)
      die("dependency cycle through " + label);
    std::string pkg, name;
    const target &t = find(label, pkg, name);
    const bool cuda = t.rule == "cuda";
    if (!cuda && t.rule != "c++")
      die(label + ": unknown rule `" + t.rule + "`");
    const fs::path pdir = root / pkg, out = pdir / ".out";
    if (!dry_run)
      fs::create_directories(out);

    product res;
    for (const auto &d : t.deps) {
      const product &p = build(d);
      res.objects.insert(res.objects.end(), p.objects.begin(), p.objects.end());
      res.libraries.insert(res.libraries.end(), p.libraries.begin(), p.libraries.end());
      res.newest = std::max(res.newest, p.newest);
    }
    auto newest = res.newest;
    for (const auto &h : t.hdrs) {
      if (!fs::exists(pdir / h))
        die(label + ": header " + (pdir / h).string() + " does not exist");
      newest = std::max(newest, mtime(pdir / h));
    }

    std::vector<std::string> mine;
    for (const auto &s : t.srcs) {
      const fs::path src = pdir / s, obj = out / fs::path(s).replace_extension(".o");
      if (!fs::exists(src))
        die(label + ": source " + src.string() + " does not exist");
      if (force || mtime(obj) < std::max(mtime(src), newest)) {
        std::vector<std::string> cmd;
        if (cuda)
          cmd = {nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
                 "-Xcompiler", "-fPIC"};
        else
          cmd = {cxx, "-O3", "-std=c++20", "-Wall", "-I" + (root / "dependence_free_rl_b200" / "host").string(),
                 "-I" + (root / "include").string()};
        cmd.insert(cmd.end(), t.gopts.begin(), t.gopts.end());
        cmd.insert(cmd.end(), {"-c", src.string(), "-o", obj.string()});
        run(cmd);
      }
      mine.push_back(obj.string());
      newest = std::max(newest, mtime(obj));
    }

    const bool shared = std::find(t.lopts.begin(), t.lopts.end(), "-shared") != t.lopts.end();
    if (cuda && shared) {
      const fs::path lib = out / ("lib" + name + ".so");
      if (force || mtime(lib) < newest) {
        std::vector<std::string> cmd = {nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", lib.string()};
        cmd.insert(cmd.end(), mine.begin(), mine.end());
        cmd.push_back("-ldl");
        run(cmd);
      }
      res.objects.clear();
      res.libraries.push_back(lib.string());
      res.newest = std::max(newest, mtime(lib));
    } else if (t.main) {
      const fs::path exe = out / name;
      if (force || mtime(exe) < newest) {
        std::vector<std::string> cmd = {cxx};
        cmd.insert(cmd.end(), mine.begin(), mine.end());
        cmd.insert(cmd.end(), res.objects.begin(), res.objects.end());
        std::set<std::string> seen;
        for (const auto &l : res.libraries) {
          if (!seen.insert(l).second)
            continue;
          const fs::path lp(l);
          const std::string stem = lp.stem().string();  // lib<name>
          cmd.insert(cmd.end(), {"-L" + lp.parent_path().string(), "-l" + stem.substr(3),
                                 "-Wl,-rpath," + lp.parent_path().string()});
        }
        cmd.insert(cmd.end(), t.lopts.begin(), t.lopts.end());
        cmd.insert(cmd.end(), {"-pthread", "-o", exe.string()});
        run(cmd);
      }
      res = product{};
      res.newest = mtime(exe);
    } else {  // header-only / object target: hands its objects and libraries to the dependents
      res.objects.insert(res.objects.end(), mine.begin(), mine.end());
      res.newest = newest;
    }
    open.erase(label);
    return done[label] = res;
  }
};

fs::path find_root(fs::path dir) {
  for (;; dir = dir.parent_path()) {
    if (fs::exists(dir / "include" / "dfrl.h"))
      return dir;
    if (dir == dir.root_path())
      die("no repo root (a directory holding include/dfrl.h) above the current directory");
  }
}

}  // namespace

int main(int argc, char **argv) {
  builder b;
  if (const char *e = std::getenv("CXX"))
    b.cxx = e;
  if (const char *e = std::getenv("NVCC"))
    b.nvcc = e;
  std::vector<std::string> labels;
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    if (a == "-n")
      b.dry_run = true;
    else if (a == "-B")
      b.force = true;
    else if (a == "-h" || a == "--help")
      labels.clear(), i = argc;
    else
      labels.push_back(a);
  }
  if (labels.empty()) {
    std::fprintf(stderr, "usage: xmake [-n] [-B] <target | //pkg/target> ...\n");
    return 2;
  }
  const fs::path cwd = fs::current_path();
  b.root = find_root(cwd);
  for (std::string label : labels) {
    if (label.rfind("//", 0) != 0)  // bare name: a target of the package in the current directory
      label = "//" + fs::relative(cwd, b.root).generic_string() + "/" + label;
    b.build(label);
  }
  return 0;
}
