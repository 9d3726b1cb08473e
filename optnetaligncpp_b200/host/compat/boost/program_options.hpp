// Stand-in for <boost/program_options.hpp>, which is absent in this image.  It covers exactly the
// subset the reference's argParsing.h / drivers use (argParsing.h:36-151, steadystate.cpp:42-72):
// options_description + add_options()(name, description) / (name, value<T>(), description),
// variables_map with count() / operator[] / as<T>(), parse_command_line, store, notify and
// operator<< for --help.  Long options only ("--name", "--name value", "--name=value"), like the
// reference's own invocations (manual.tex:131-142).  With a real Boost on the include path this
// directory is simply left out (see ../Makefile).
#pragma once
#include <iomanip>
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

// a parsed value of one of the types the reference asks for (string / int / float)
class variable_value {
public:
    variable_value() {}
    explicit variable_value(const std::string& text, const std::string& type) : text_(text), type_(type) {}
    template <typename T>
    T as() const {
        if (type_ != typeid(T).name()) throw error("option value requested with the wrong type");
        T v;
        parse(v);
        return v;
    }
    bool empty() const { return type_.empty(); }

private:
    void parse(std::string& v) const { v = text_; }
    template <typename T>
    void parse(T& v) const {
        std::istringstream in(text_);
        in >> v;
    }
    std::string text_, type_;
};

class value_semantic {
public:
    virtual ~value_semantic() {}
    virtual std::string type() const = 0;
    virtual bool valid(const std::string& text) const = 0;
};
template <typename T>
class typed_value : public value_semantic {
public:
    std::string type() const override { return typeid(T).name(); }
    bool valid(const std::string& text) const override { return check(text, static_cast<T*>(nullptr)); }

private:
    static bool check(const std::string&, std::string*) { return true; }
    template <typename U>
    static bool check(const std::string& text, U*) {
        std::istringstream in(text);
        U v;
        in >> v;
        return !in.fail() && in.eof();
    }
};
template <typename T>
typed_value<T>* value() { return new typed_value<T>(); }

struct option_description {
    std::string name, description;
    std::shared_ptr<const value_semantic> semantic;   // null: a flag without an argument
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
    explicit options_description(const std::string& caption = "") : caption_(caption) {}
    options_description_easy_init add_options() { return options_description_easy_init(this); }
    const option_description* find(const std::string& name) const {
        for (const auto& o : options_)
            if (o.name == name) return &o;
        return nullptr;
    }
    void add(const option_description& o) { options_.push_back(o); }
    friend std::ostream& operator<<(std::ostream& os, const options_description& d) {
        if (!d.caption_.empty()) os << d.caption_ << ":\n";
        for (const auto& o : d.options_) {
            std::string left = "  --" + o.name + (o.semantic ? " arg" : "");
            os << std::left << std::setw(26) << left << " " << o.description << "\n";
        }
        return os;
    }

private:
    std::string caption_;
    std::vector<option_description> options_;
};
inline options_description_easy_init& options_description_easy_init::operator()(const char* name,
                                                                                const char* description) {
    owner_->add(option_description{name, description, nullptr});
    return *this;
}
inline options_description_easy_init& options_description_easy_init::operator()(const char* name,
                                                                                const value_semantic* s,
                                                                                const char* description) {
    owner_->add(option_description{name, description, std::shared_ptr<const value_semantic>(s)});
    return *this;
}

struct parsed_option {
    std::string name, text, type;
};
typedef std::vector<parsed_option> parsed_options;

inline parsed_options parse_command_line(int argc, const char* const argv[], const options_description& desc) {
    parsed_options out;
    for (int i = 1; i < argc; ++i) {
        std::string tok = argv[i];
        if (tok.size() < 3 || tok.compare(0, 2, "--") != 0)
            throw error("too many positional options have been specified on the command line");
        std::string name = tok.substr(2), text;
        bool has_text = false;
        const size_t eq = name.find('=');
        if (eq != std::string::npos) {
            text = name.substr(eq + 1);
            name = name.substr(0, eq);
            has_text = true;
        }
        const option_description* o = desc.find(name);
        if (!o) throw error("unrecognised option '--" + name + "'");
        if (!o->semantic) {
            if (has_text) throw error("option '--" + name + "' does not take any arguments");
            out.push_back(parsed_option{name, "", ""});
            continue;
        }
        if (!has_text) {
            if (i + 1 >= argc) throw error("the required argument for option '--" + name + "' is missing");
            text = argv[++i];
        }
        if (!o->semantic->valid(text))
            throw error("the argument ('" + text + "') for option '--" + name + "' is invalid");
        out.push_back(parsed_option{name, text, o->semantic->type()});
    }
    return out;
}

class variables_map : public std::map<std::string, variable_value> {
public:
    const variable_value& operator[](const std::string& name) const {
        static const variable_value empty;
        const_iterator it = find(name);
        return it == end() ? empty : it->second;
    }
    variable_value& slot(const std::string& name) { return std::map<std::string, variable_value>::operator[](name); }
};

inline void store(const parsed_options& parsed, variables_map& vm) {
    for (const auto& p : parsed)
        if (!vm.count(p.name)) vm.slot(p.name) = variable_value(p.text, p.type.empty() ? "flag" : p.type);   // first occurrence wins
}
inline void notify(variables_map&) {}

}  // namespace program_options
}  // namespace boost
