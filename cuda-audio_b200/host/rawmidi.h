// rawmidi.h -- MIDI control surface of the convolution engine.
//
// Same surface as the reference's MIDI layer (src/midi.h:10-46): `RawMidi::MessageHandler`
// (Convolution implements it, conv.h:30), `RawMidi::Device` with id / handler / start() / stop() /
// send() / isOpen / isRunning, so main.cu:43-52,86-87 ports line by line.  Differences, on purpose:
//  * no libasound: the device id "hw:C,D[,S]" is opened as the kernel's raw MIDI character device
//    /dev/snd/midiC<C>D<D> (what snd_rawmidi_open ends up reading); any other id is taken as a path
//    (a FIFO or file in the tests, /dev/midi1 on OSS-style systems);
//  * the reader blocks in poll() instead of spinning with usleep(1000) (midi.cu:43-47);
//  * `MidiParser` is a complete running-status stream parser: 2-byte messages (program change,
//    channel pressure), system common and real-time bytes are handled instead of asserting
//    (midi.cu:16-18), SysEx is delivered whole when it fits 256 bytes and dropped otherwise.
#pragma once
#include <atomic>
#include <cstddef>
#include <cstdint>
#include <functional>
#include <string>
#include <thread>

// Byte stream -> complete MIDI messages (status byte first, running status expanded).
class MidiParser {
public:
    using Sink = std::function<void(const uint8_t *msg, size_t len)>;
    explicit MidiParser(Sink sink) : _sink(std::move(sink)) {}
    void feed(const uint8_t *bytes, size_t n);
    void reset() { _len = 0; _need = 0; _running = 0; _sysex = false; _overflow = false; }

private:
    static int dataBytes(uint8_t status);
    Sink _sink;
    uint8_t _buf[256];
    size_t _len = 0, _need = 0;
    uint8_t _running = 0;
    bool _sysex = false, _overflow = false;
};

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
        virtual ~Device() { if (isOpen) stop(); }
        Device(const Device &) = delete;
        Device &operator=(const Device &) = delete;

        // opens the device and starts the reader thread; false (with `error` set) if it cannot be opened
        bool start();
        void stop();
        bool send(const uint8_t *data, size_t len);
        // deliver one complete MIDI message to the handler (what the reader thread does per message)
        void inject(const uint8_t *data, size_t len) const { if (handler) handler->onMidiMessage(this, data, len); }
        // "hw:2,0" -> "/dev/snd/midiC2D0"; anything else is returned unchanged
        static std::string devicePath(const std::string &id);

        MessageHandler *handler = nullptr;
        std::string id, error;
        bool isOpen = false;
        std::atomic<bool> isRunning{false};

    private:
        void run();
        int _fd = -1;
        bool _writable = false;
        std::thread _thread;
    };
};
