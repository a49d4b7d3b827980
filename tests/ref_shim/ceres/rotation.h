// stand-in: nothing of ceres/rotation.h is used by the translation units built into oracle/_ref
