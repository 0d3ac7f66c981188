#pragma once
#include <boost/version.hpp>
#define BOOST_CLASS_VERSION(T, N)
namespace boost { namespace serialization { class access; } }
