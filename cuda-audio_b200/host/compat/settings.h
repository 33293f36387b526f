/* source-compatibility shim: code written against limitz/cuda-audio src/settings.h compiles against the B200 engine */
#pragma once
#include "../settings_file.h"
