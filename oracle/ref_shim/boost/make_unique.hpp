#pragma once
#include <memory>
namespace boost { using std::make_unique; }
