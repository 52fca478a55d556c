"""B200-native hot path of kaldi-ctc: CTC loss+gradient and the LSTM/GRU
recurrent forward/backward, behind the reference's own interfaces.

The compute lives in csrc/ (hand-written CUDA for sm_100a) behind the C ABI
declared in include/ctc.h and include/b200rnn.h; this package is the thin
host-side mirror used by tests and bench.py.  There is no CPU fallback: every
compute entry point raises if libb200ctc.so is missing or no B200 is present.
"""
__version__ = "0.1.0"
