/*
 * Minimal stand-in for <boost/program_options.hpp>: exactly the subset the
 * reference host program uses (main.cpp:58-78, main_aux_functions.h:77-145).
 * Written for this repo because the image ships no boost headers.  TEST
 * INFRASTRUCTURE: only used to build oracle/_ref (the unmodified reference).
 */
#ifndef ORACLE_SHIM_BOOST_PO_HPP
#define ORACLE_SHIM_BOOST_PO_HPP
#include <math.h>
#include <algorithm>
#include <cmath>
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace boost {
namespace program_options {

struct value_base {
    bool has_default = false;
    virtual ~value_base() {}
    virtual bool takes_arg() const = 0;
    virtual void parse(const std::string &s) = 0;  // parse text into the holder
    virtual void apply_default() = 0;
    virtual void notify() = 0;                      // write to the bound variable
    virtual const void *get(const std::type_info &) const = 0;
};

template <typename T>
struct typed_value : value_base {
    T *target;
    T val{}, def{};
    explicit typed_value(T *t) : target(t) {}
    typed_value *default_value(const T &d) {
        def = d;
        has_default = true;
        return this;
    }
    bool takes_arg() const override { return true; }
    void parse(const std::string &s) override {
        std::istringstream is(s);
        if (!(is >> val)) throw std::runtime_error("the argument ('" + s + "') for an option is invalid");
    }
    void apply_default() override { val = def; }
    void notify() override {
        if (target) *target = val;
    }
    const void *get(const std::type_info &) const override { return &val; }
};
template <>
inline void typed_value<std::string>::parse(const std::string &s) { val = s; }

struct flag_value : value_base {
    bool takes_arg() const override { return false; }
    void parse(const std::string &) override {}
    void apply_default() override {}
    void notify() override {}
    const void *get(const std::type_info &) const override { return nullptr; }
};

template <typename T>
typed_value<T> *value(T *t) { return new typed_value<T>(t); }

struct option_desc {
    std::string long_name, short_name, help;
    std::shared_ptr<value_base> sem;
};

class options_description;
class easy_init {
    options_description *owner;
public:
    explicit easy_init(options_description *o) : owner(o) {}
    easy_init &operator()(const char *name, const char *help);
    easy_init &operator()(const char *name, value_base *sem, const char *help);
};

class options_description {
public:
    std::string caption;
    std::vector<option_desc> opts;
    explicit options_description(const std::string &c) : caption(c) {}
    easy_init add_options() { return easy_init(this); }
    void add(const char *name, value_base *sem, const char *help) {
        option_desc d;
        std::string n(name);
        size_t comma = n.find(',');
        d.long_name = n.substr(0, comma);
        if (comma != std::string::npos) d.short_name = n.substr(comma + 1);
        d.help = help;
        d.sem.reset(sem);
        opts.push_back(d);
    }
    const option_desc *find_long(const std::string &n) const {
        for (auto &o : opts) if (o.long_name == n) return &o;
        return nullptr;
    }
    const option_desc *find_short(const std::string &n) const {
        for (auto &o : opts) if (!o.short_name.empty() && o.short_name == n) return &o;
        return nullptr;
    }
};
inline easy_init &easy_init::operator()(const char *name, const char *help) {
    owner->add(name, new flag_value(), help);
    return *this;
}
inline easy_init &easy_init::operator()(const char *name, value_base *sem, const char *help) {
    owner->add(name, sem, help);
    return *this;
}
inline std::ostream &operator<<(std::ostream &os, const options_description &d) {
    os << d.caption << ":\n";
    for (auto &o : d.opts) {
        std::string left = "  ";
        if (!o.short_name.empty()) left += "-" + o.short_name + " [ --" + o.long_name + " ]";
        else left += "--" + o.long_name;
        if (o.sem->takes_arg()) left += " arg";
        os << left << "  " << o.help << "\n";
    }
    return os;
}

class variable_value {
public:
    std::shared_ptr<value_base> sem;
    bool is_defaulted = false, is_empty = true;
    bool defaulted() const { return is_defaulted; }
    bool empty() const { return is_empty; }
    template <typename T>
    const T &as() const {
        if (is_empty) throw std::runtime_error("boost::bad_any_cast: failed conversion using boost::any_cast");
        return *static_cast<const T *>(sem->get(typeid(T)));
    }
};

class variables_map : public std::map<std::string, variable_value> {
public:
    const variable_value &operator[](const std::string &k) const {
        static const variable_value none;
        auto it = find(k);
        return it == end() ? none : it->second;
    }
    variable_value &slot(const std::string &k) { return std::map<std::string, variable_value>::operator[](k); }
};

struct parsed_options {
    const options_description *desc;
    std::vector<std::pair<const option_desc *, std::string>> items;
};

inline parsed_options parse_command_line(int argc, const char *const *argv, const options_description &desc) {
    parsed_options p;
    p.desc = &desc;
    for (int i = 1; i < argc; i++) {
        std::string a(argv[i]);
        const option_desc *o = nullptr;
        std::string val;
        bool have_val = false;
        if (a.size() > 2 && a[0] == '-' && a[1] == '-') {
            std::string n = a.substr(2);
            size_t eq = n.find('=');
            if (eq != std::string::npos) { val = n.substr(eq + 1); n = n.substr(0, eq); have_val = true; }
            o = desc.find_long(n);
            if (!o) throw std::runtime_error("unrecognised option '" + a + "'");
        } else if (a.size() >= 2 && a[0] == '-') {
            o = desc.find_short(a.substr(1, 1));
            if (!o) throw std::runtime_error("unrecognised option '" + a + "'");
            if (a.size() > 2) { val = a.substr(2); have_val = true; }
        } else {
            throw std::runtime_error("too many positional options have been specified on the command line");
        }
        if (o->sem->takes_arg() && !have_val) {
            if (i + 1 >= argc) throw std::runtime_error("the required argument for option '" + a + "' is missing");
            val = argv[++i];
        }
        p.items.push_back({o, val});
    }
    return p;
}

inline void store(const parsed_options &p, variables_map &vm) {
    for (auto &it : p.items) {
        variable_value &v = vm.slot(it.first->long_name);
        v.sem = it.first->sem;
        if (v.sem->takes_arg()) v.sem->parse(it.second);
        v.is_empty = false;
        v.is_defaulted = false;
    }
    for (auto &o : p.desc->opts) {
        if (vm.count(o.long_name)) continue;
        if (o.sem->has_default) {
            variable_value &v = vm.slot(o.long_name);
            v.sem = o.sem;
            v.sem->apply_default();
            v.is_empty = false;
            v.is_defaulted = true;
        }
    }
}

inline void notify(variables_map &vm) {
    for (auto &kv : vm) if (kv.second.sem) kv.second.sem->notify();
}

}  // namespace program_options
}  // namespace boost
#endif
