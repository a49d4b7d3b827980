// stand-in: Pangolin is only used by the viewer (out of scope); Camera.h includes it unconditionally
