// TEST INFRASTRUCTURE — not TensorFlow (see op.h in this directory).
#pragma once
#include "tensorflow/core/framework/op.h"
