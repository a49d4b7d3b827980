// stand-in for boost::bind with the global _1, _2 placeholders (ref: src/Feature_alignment.cpp:88)
#ifndef MINI_BOOST_BIND_H
#define MINI_BOOST_BIND_H
#include <functional>
namespace boost { using std::bind; }
using std::placeholders::_1;
using std::placeholders::_2;
#endif
