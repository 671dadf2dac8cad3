"""Drop-in for the two ``torch_geometric`` symbols the reference imports
(/root/reference/model/layers.py:7-8, graphs/graph.py:4, model/evaluation.py:5), backed by the
B200 engine.  Put ``scaling-rgcn-training_b200/`` on sys.path ahead of any real PyG and the
reference's model/, graphs/ files run unchanged on the GPU."""
__version__ = '2.3.1+rgcn_b200'
