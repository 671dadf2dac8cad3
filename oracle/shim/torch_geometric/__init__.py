"""ORACLE-BACKED stand-in for the two torch_geometric symbols the reference imports
(model/layers.py:7-8, graphs/graph.py:4, model/evaluation.py:5).  Test infrastructure:
lets the UNMODIFIED reference callers run on CPU on top of oracle/rgcn_oracle.py.
The product's CUDA-only drop-in lives in scaling-rgcn-training_b200/torch_geometric/."""
__version__ = '2.3.1+oracle'
