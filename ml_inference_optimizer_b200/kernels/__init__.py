"""Host-side mirror of the reference's ``kernels`` package (attention, mlp and the functional ``triton`` entry points),
executing on the sm_100a C-ABI library. No Triton is involved; the module names are kept so imports stay drop-in."""
