"""gtn_applications_b200 — the WFST sequence-criterion hot path of
facebookresearch/gtn_applications (criterions/{ctc,asg,stc,transducer}.py over the
external GTN library) rebuilt for B200: hand-written sm_100a CUDA kernels in
libwfst_b200.so behind a C ABI (include/wfst_b200.h), exposed through
torch.autograd.Function classes with the reference's signatures.

There is no CPU fallback: every compute call needs the CUDA library and a CUDA
device and raises otherwise."""

__version__ = "0.1.0"
