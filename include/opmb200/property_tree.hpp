// property_tree.hpp -- dependency-free stand-in for Opm::PropertyTree
// (opm/simulators/linalg/PropertyTree.{hpp,cpp}: a thin wrapper over boost::property_tree::ptree
// + read_json).  Same surface: get<T>(key[, default]), put, get_child, get_child_optional,
// get_child_keys, JSON in / JSON out.  Like the boost tree every leaf is stored as a string
// ("tol": "0.5" and "tol": 0.5 read the same) and keys may be dotted paths
// ("preconditioner.type").
#pragma once
#include <cctype>
#include <cstdlib>
#include <fstream>
#include <map>
#include <memory>
#include <optional>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace opmb200 {

class PropertyTree
{
public:
    PropertyTree() = default;

    // Opm::PropertyTree(const std::string& jsonFile)
    static PropertyTree fromFile(const std::string& path)
    {
        std::ifstream f(path);
        if (!f)
            throw std::invalid_argument("PropertyTree: cannot open " + path);
        std::stringstream ss;
        ss << f.rdbuf();
        return fromJson(ss.str());
    }

    static PropertyTree fromJson(const std::string& text)
    {
        Parser p {text, 0};
        p.skip();
        PropertyTree t = p.value();
        p.skip();
        if (p.pos != text.size())
            throw std::invalid_argument("PropertyTree: trailing characters in JSON");
        if (t.isLeaf_ && !text.empty())
            throw std::invalid_argument("PropertyTree: top level must be an object");
        return t;
    }

    template <class T>
    T get(const std::string& key) const
    {
        const PropertyTree* n = find(key);
        if (!n || !n->isLeaf_)
            throw std::invalid_argument("PropertyTree: no such key: " + key);
        return convert<T>(n->value_, key);
    }

    template <class T>
    T get(const std::string& key, const T& defValue) const
    {
        const PropertyTree* n = find(key);
        if (!n || !n->isLeaf_)
            return defValue;
        return convert<T>(n->value_, key);
    }

    template <class T>
    void put(const std::string& key, const T& value)
    {
        std::ostringstream os;
        os.precision(17);
        os << value;
        PropertyTree* n = this;
        size_t start = 0;
        while (true) {
            const size_t dot = key.find('.', start);
            const std::string part = key.substr(start, dot == std::string::npos ? dot : dot - start);
            n->isLeaf_ = false;
            n = &n->child(part);
            if (dot == std::string::npos)
                break;
            start = dot + 1;
        }
        n->isLeaf_ = true;
        n->value_ = os.str();
        n->children_.clear();
    }

    PropertyTree get_child(const std::string& key) const
    {
        const PropertyTree* n = find(key);
        if (!n)
            throw std::invalid_argument("PropertyTree: no such child: " + key);
        return *n;
    }

    std::optional<PropertyTree> get_child_optional(const std::string& key) const
    {
        const PropertyTree* n = find(key);
        if (!n)
            return std::nullopt;
        return *n;
    }

    std::vector<std::string> get_child_keys() const
    {
        std::vector<std::string> k;
        for (const auto& c : children_)
            k.push_back(c.first);
        return k;
    }

    bool empty() const { return children_.empty() && value_.empty(); }

    // Opm::PropertyTree::write_json
    std::string toJson(int indent = 0) const
    {
        if (isLeaf_)
            return quote(value_);
        std::string pad(indent + 4, ' '), out = "{\n";
        for (size_t i = 0; i < children_.size(); ++i) {
            out += pad + quote(children_[i].first) + ": " + children_[i].second->toJson(indent + 4);
            out += (i + 1 < children_.size()) ? ",\n" : "\n";
        }
        return out + std::string(indent, ' ') + "}";
    }

private:
    bool isLeaf_ = false;
    std::string value_;
    std::vector<std::pair<std::string, std::shared_ptr<PropertyTree>>> children_; // insertion order

    PropertyTree& child(const std::string& name)
    {
        for (auto& c : children_)
            if (c.first == name)
                return *c.second;
        children_.emplace_back(name, std::make_shared<PropertyTree>());
        return *children_.back().second;
    }

    const PropertyTree* find(const std::string& key) const
    {
        const PropertyTree* n = this;
        size_t start = 0;
        while (true) {
            const size_t dot = key.find('.', start);
            const std::string part = key.substr(start, dot == std::string::npos ? dot : dot - start);
            const PropertyTree* next = nullptr;
            for (const auto& c : n->children_)
                if (c.first == part)
                    next = c.second.get();
            if (!next)
                return nullptr;
            n = next;
            if (dot == std::string::npos)
                return n;
            start = dot + 1;
        }
    }

    static std::string quote(const std::string& s)
    {
        std::string o = "\"";
        for (char c : s) {
            if (c == '"' || c == '\\')
                o += '\\';
            o += c;
        }
        return o + "\"";
    }

    template <class T>
    static T convert(const std::string& s, const std::string& key)
    {
        if constexpr (std::is_same_v<T, std::string>) {
            return s;
        } else if constexpr (std::is_same_v<T, bool>) {
            if (s == "true" || s == "1")
                return true;
            if (s == "false" || s == "0")
                return false;
            throw std::invalid_argument("PropertyTree: key " + key + " is not a bool: " + s);
        } else {
            std::istringstream is(s);
            T v {};
            is >> v;
            if (is.fail() || !(is >> std::ws).eof())
                throw std::invalid_argument("PropertyTree: cannot convert value of " + key + ": " + s);
            return v;
        }
    }

    struct Parser {
        const std::string& t;
        size_t pos;
        void skip()
        {
            while (pos < t.size() && std::isspace((unsigned char)t[pos]))
                ++pos;
        }
        [[noreturn]] void fail(const char* what) const
        {
            throw std::invalid_argument(std::string("PropertyTree: JSON parse error (") + what + ") at offset "
                                        + std::to_string(pos));
        }
        std::string str()
        {
            std::string o;
            ++pos; // opening quote
            while (pos < t.size() && t[pos] != '"') {
                if (t[pos] == '\\' && pos + 1 < t.size()) {
                    ++pos;
                    switch (t[pos]) {
                    case 'n': o += '\n'; break;
                    case 't': o += '\t'; break;
                    default: o += t[pos];
                    }
                } else {
                    o += t[pos];
                }
                ++pos;
            }
            if (pos >= t.size())
                fail("unterminated string");
            ++pos;
            return o;
        }
        PropertyTree value()
        {
            skip();
            if (pos >= t.size())
                fail("unexpected end");
            PropertyTree n;
            if (t[pos] == '{') {
                ++pos;
                skip();
                if (pos < t.size() && t[pos] == '}') {
                    ++pos;
                    return n;
                }
                while (true) {
                    skip();
                    if (pos >= t.size() || t[pos] != '"')
                        fail("expected key");
                    const std::string key = str();
                    skip();
                    if (pos >= t.size() || t[pos] != ':')
                        fail("expected ':'");
                    ++pos;
                    n.child(key) = value();
                    skip();
                    if (pos < t.size() && t[pos] == ',') {
                        ++pos;
                        continue;
                    }
                    if (pos < t.size() && t[pos] == '}') {
                        ++pos;
                        break;
                    }
                    fail("expected ',' or '}'");
                }
                return n;
            }
            if (t[pos] == '[') { // arrays become children with empty keys, like boost's read_json
                ++pos;
                skip();
                if (pos < t.size() && t[pos] == ']') {
                    ++pos;
                    return n;
                }
                while (true) {
                    n.children_.emplace_back("", std::make_shared<PropertyTree>(value()));
                    skip();
                    if (pos < t.size() && t[pos] == ',') {
                        ++pos;
                        continue;
                    }
                    if (pos < t.size() && t[pos] == ']') {
                        ++pos;
                        break;
                    }
                    fail("expected ',' or ']'");
                }
                return n;
            }
            n.isLeaf_ = true;
            if (t[pos] == '"') {
                n.value_ = str();
                return n;
            }
            const size_t s = pos;
            while (pos < t.size() && (std::isalnum((unsigned char)t[pos]) || t[pos] == '.' || t[pos] == '-' || t[pos] == '+'))
                ++pos;
            if (pos == s)
                fail("unexpected character");
            n.value_ = t.substr(s, pos - s);
            return n;
        }
    };
};

} // namespace opmb200
