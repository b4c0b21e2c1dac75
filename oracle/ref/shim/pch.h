// TEST INFRASTRUCTURE (oracle/_ref build). Stand-in for the reference's precompiled
// header (/root/reference/pch/pch.h:25-50), which drags in Boost headers that are not
// installed here. The hot-path translation units use no Boost symbol, so the standard
// headers below are all they need.
#pragma once
#include <iostream>
#include <fstream>
#include <sstream>
#include <functional>
#include <thread>
#include <string>
#include <atomic>
#include <mutex>
#include <random>
#include <unordered_map>
#include <unordered_set>
#include <chrono>
#include <regex>
#include <tuple>
#include <numeric>
#include <algorithm>
#include <memory>
#include <cmath>
#include <cstring>
#include <cassert>
#include <vector>
#include <map>
#include <lightmetrica/macros.h>

// trianglemesh_obj.cpp:48-50 joins the property tree's base path and the `path` parameter with boost::filesystem::path
// and turns the result back into a string: the only Boost use in the translation units compiled here.
namespace boost { namespace filesystem {
class path {
public:
    path() {}
    path(const std::string& s) : s_(s) {}
    path(const char* s) : s_(s) {}
    path operator/(const path& rhs) const
    {
        if (s_.empty() || (!rhs.s_.empty() && rhs.s_[0] == '/')) return rhs;
        return path(s_.back() == '/' ? s_ + rhs.s_ : s_ + "/" + rhs.s_);
    }
    const std::string& string() const { return s_; }
private:
    std::string s_;
};
} }
