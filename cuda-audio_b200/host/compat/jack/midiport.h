/* nothing from <jack/midiport.h> is used on this path */
