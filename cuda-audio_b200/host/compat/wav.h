/* source-compatibility shim: code written against limitz/cuda-audio src/wav.h compiles against the B200 engine */
#pragma once
#include "../wavfile.h"
