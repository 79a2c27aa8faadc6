// include/df/df.hpp -- stands where the reference's digital-filtering-c++/df/df.hpp stands, so that a caller written against the
// reference (its own test/cpp-main.cpp: `#include "../df/df.hpp"`) compiles UNCHANGED against the B200 library: copy or link this
// directory next to the caller's test/ directory.  Like the reference's header it pulls the std names in.
#pragma once
#include <cmath>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>
#include "../digital_filter.hpp"
using namespace std;      // df.hpp:18
