// stand-in: Pangolin is only used by the viewer (out of scope); include/Camera.h includes it unconditionally and include/Viewer.h
// names one of its types in two declarations
#ifndef MINI_PANGOLIN_H
#define MINI_PANGOLIN_H
#include <numeric>
#include <string>
#include <vector>
// src/Tracking.cpp writes `vector`, `string`, `endl` unqualified inside namespace DSDTM: one of the real third-party headers leaks
// `using namespace std;` into the global namespace. The stand-in does the same so that the file compiles unmodified.
using namespace std;
namespace pangolin {
struct OpenGlMatrix { double m[16]; };
}
#endif
