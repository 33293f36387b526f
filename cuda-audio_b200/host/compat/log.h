/* source-compatibility shim: code written against limitz/cuda-audio src/log.h compiles against the B200 engine */
#pragma once
#include "../logger.h"
