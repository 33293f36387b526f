// rawmidi.h -- the two types of the reference's MIDI layer (src/midi.h:10-46) that the
// Convolution class is written against: the message-handler interface and the device handle
// CC maps point at.  The ALSA rawmidi reader thread (midi.cu) is a control-plane component and
// out of scope; any thread may call handler->onMidiMessage().
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>

class RawMidi {
public:
    class Device;
    class MessageHandler {
    public:
        virtual ~MessageHandler() = default;
        virtual void onMidiMessage(const Device *sender, const uint8_t *buffer, size_t len) = 0;
    };
    class Device {
    public:
        explicit Device(const std::string &id) : id(id) {}
        virtual ~Device() = default;
        // deliver one complete MIDI message (what midi.cu's reader thread does, midi.cu:22-59)
        void inject(const uint8_t *data, size_t len) const { if (handler) handler->onMidiMessage(this, data, len); }
        MessageHandler *handler = nullptr;
        std::string id;
        bool isOpen = false, isRunning = false;
    };
};
