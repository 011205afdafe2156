// Minimal stand-in for Boost.program_options -- TEST SCAFFOLDING ONLY (Boost is absent from this image).
// Implements exactly the subset the reference's example/main.cpp:55-93 uses, so that file can be compiled
// UNCHANGED against the drop-in headers of this repository (tests/test_cpp_dropin.py).
#pragma once
#include <any>
#include <iostream>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace boost { namespace program_options {

class error : public std::logic_error { public: using std::logic_error::logic_error; };

struct value_semantic {
    bool has_default = false, is_required = false;
    std::any default_any;
    virtual ~value_semantic() = default;
    virtual std::any parse(const std::string &s) const = 0;
};
template <typename T> struct typed_value : value_semantic {
    typed_value *default_value(const T &v) { has_default = true; default_any = v; return this; }
    typed_value *required() { is_required = true; return this; }
    std::any parse(const std::string &s) const override {
        if constexpr (std::is_same_v<T, std::string>) return s;
        else { try { return (T)std::stoll(s); } catch (...) { throw error("the argument ('" + s + "') is invalid"); } }
    }
};
template <typename T> typed_value<T> *value() { return new typed_value<T>(); }

struct option_def { std::string long_name, short_name, help; std::shared_ptr<value_semantic> sem; };

class options_description;
class options_easy_init {
    options_description *owner;
public:
    explicit options_easy_init(options_description *o) : owner(o) {}
    options_easy_init &operator()(const char *name, const char *help);
    options_easy_init &operator()(const char *name, value_semantic *sem, const char *help);
};
class options_description {
public:
    std::string caption; std::vector<option_def> opts;
    explicit options_description(const std::string &c) : caption(c) {}
    options_easy_init add_options() { return options_easy_init(this); }
    void add(const char *name, value_semantic *sem, const char *help) {
        std::string n(name); option_def d; auto c = n.find(',');
        d.long_name = n.substr(0, c); if (c != std::string::npos) d.short_name = n.substr(c + 1);
        d.help = help; d.sem.reset(sem); opts.push_back(d);
    }
    const option_def *find(const std::string &tok) const {
        for (auto &o : opts) if (tok == "--" + o.long_name || (!o.short_name.empty() && tok == "-" + o.short_name)) return &o;
        return nullptr;
    }
};
inline options_easy_init &options_easy_init::operator()(const char *n, const char *h) { owner->add(n, nullptr, h); return *this; }
inline options_easy_init &options_easy_init::operator()(const char *n, value_semantic *s, const char *h) { owner->add(n, s, h); return *this; }
inline std::ostream &operator<<(std::ostream &os, const options_description &d) {
    os << d.caption << ":\n"; for (auto &o : d.opts) os << "  --" << o.long_name << "  " << o.help << "\n"; return os;
}

struct variable_value { std::any v; template <typename T> const T &as() const { return *std::any_cast<T>(&v); } };
class variables_map : public std::map<std::string, variable_value> {
public:
    std::vector<std::string> missing_required;
    size_t count(const std::string &k) const { return std::map<std::string, variable_value>::count(k); }
    const variable_value &operator[](const std::string &k) const { return at(k); }
};
struct parsed_options { const options_description *desc; std::vector<std::pair<std::string, std::string>> kv; };

inline parsed_options parse_command_line(int argc, char **argv, const options_description &desc) {
    parsed_options p{&desc, {}};
    for (int i = 1; i < argc; i++) {
        const option_def *o = desc.find(argv[i]);
        if (!o) throw error(std::string("unrecognised option '") + argv[i] + "'");
        if (o->sem) { if (i + 1 >= argc) throw error("the required argument for option '--" + o->long_name + "' is missing");
                      p.kv.push_back({o->long_name, argv[++i]}); }
        else p.kv.push_back({o->long_name, ""});
    }
    return p;
}
inline void store(const parsed_options &p, variables_map &vm) {
    for (auto &kv : p.kv) {
        const option_def *o = p.desc->find("--" + kv.first);
        variable_value vv; if (o->sem) vv.v = o->sem->parse(kv.second); vm.insert({kv.first, vv});
    }
    for (auto &o : p.desc->opts) {
        if (!o.sem || vm.count(o.long_name)) continue;
        if (o.sem->has_default) { variable_value vv; vv.v = o.sem->default_any; vm.insert({o.long_name, vv}); }
        else if (o.sem->is_required) vm.missing_required.push_back(o.long_name);
    }
}
inline void notify(variables_map &vm) {
    if (!vm.missing_required.empty()) throw error("the option '--" + vm.missing_required[0] + "' is required but missing");
}
}} // namespace boost::program_options
