/* source-compatibility shim: code written against limitz/cuda-audio src/conv.h compiles against the B200 engine */
#pragma once
#include "../convolution.h"
