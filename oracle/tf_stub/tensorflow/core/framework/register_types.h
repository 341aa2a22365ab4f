#include "tensorflow/core/framework/tensor.h"
