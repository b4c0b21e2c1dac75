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
