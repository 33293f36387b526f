/* empty stand-in for <jack/midiport.h>; the reference includes it but uses nothing from it */
#pragma once
