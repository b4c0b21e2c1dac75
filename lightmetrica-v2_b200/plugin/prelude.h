// Standard headers the reference's public headers rely on being pulled in by its precompiled
// header (/root/reference/pch/pch.h:25-38); force-included when building the plugins.
#pragma once
#include <limits>
#include <cmath>
#include <cstring>
#include <cassert>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>
#include <memory>
#include <functional>
#include <algorithm>
#include <unordered_map>
#include <mutex>
#include <atomic>
#include <thread>
#include <chrono>
#include <numeric>
#include <tuple>
