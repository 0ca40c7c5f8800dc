"""abmarl_b200 -- B200-native batched GridWorld engine behind Abmarl's GridWorld component API.

Build a simulation with the reference's classes (abmarl_b200.sim.gridworld.*, abmarl_b200.examples), hand it
to a manager (abmarl_b200.managers) and step thousands of copies in lockstep on the GPU.  The CUDA extension
(abmarl_b200/csrc/libbgw.so; `python -m abmarl_b200.csrc.build`) is required: there is no CPU fallback.
"""
__version__ = '0.1.0'
