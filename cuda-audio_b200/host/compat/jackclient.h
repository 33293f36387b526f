/* source-compatibility shim: code written against limitz/cuda-audio src/jackclient.h compiles against the B200 engine */
#pragma once
#include "../jack_client.h"
