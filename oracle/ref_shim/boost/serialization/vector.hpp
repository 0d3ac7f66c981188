#pragma once
#include <boost/serialization/version.hpp>
#ifndef BOOST_SERIALIZATION_SPLIT_MEMBER
#define BOOST_SERIALIZATION_SPLIT_MEMBER()
#endif
