// TEST INFRASTRUCTURE — not TensorFlow (see ../framework/op.h).
#pragma once
