from rgcn_b200.conv import RGCNConv  # noqa: F401
