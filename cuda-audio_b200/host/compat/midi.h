/* source-compatibility shim: code written against limitz/cuda-audio src/midi.h compiles against the B200 engine */
#pragma once
#include "../rawmidi.h"
