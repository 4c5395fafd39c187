// message_assembler.h -- host half of byte_state_machine: line buffer, ZCZC / NNNN framing,
// B1B2B3B4 extraction and the add_message hand-off (receiver/nav_b_sm.C:44-52, :56-97), driven by
// the character / line / abort event stream the demod kernel emits per channel.
#pragma once
#include <regex.h>

#include <string>
#include <vector>

namespace nvx {

struct AssembledMessage {
    int stream, freq;
    std::string bbbb, text;
};

class MessageAssembler {
  public:
    MessageAssembler();
    ~MessageAssembler();
    MessageAssembler(const MessageAssembler&) = delete;
    MessageAssembler& operator=(const MessageAssembler&) = delete;

    void resize(int channels);
    void reset();
    // feed `n` event bytes of channel `ch`; completed messages are appended to `out`
    void feed(int ch, int stream, int freq, const unsigned char* ev, size_t n, std::vector<AssembledMessage>* out);

  private:
    struct Channel {
        std::string line, text, bbbb;
        bool in_message = false;
    };
    void line_done(Channel& c, int stream, int freq, std::vector<AssembledMessage>* out);
    std::vector<Channel> ch_;
    regex_t som_, eom_;
};

}  // namespace nvx
