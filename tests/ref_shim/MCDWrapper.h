// stand-in for Thirdparty/fastMCD/include/MCDWrapper.h (IplImage-era OpenCV C API; the moving-object detector is a disabled feature
// of the reference -- its call site in Tracking is commented out, ref: src/Tracking.cpp:230 -- and out of scope).
// include/Moving_Detection.h derives Moving_Detecter from this class, so the name has to exist for include/Tracking.h to parse.
#ifndef MINI_MCDWRAPPER_H
#define MINI_MCDWRAPPER_H
class MCDWrapper {
public:
    MCDWrapper() {}
    virtual ~MCDWrapper() {}
};
#endif
