// Stand-in for <boost/version.hpp> (Boost is not installed in this image).
// TEST INFRASTRUCTURE ONLY: lets the reference's own hot-path sources compile unmodified
// into oracle/_ref/. Declares third-party names only; contains no fuzzy-match logic.
#pragma once
#define BOOST_VERSION 108300
#include <memory>
#include <algorithm>
#include <cstring>
#include <iostream>
#include <string>
#include <vector>
