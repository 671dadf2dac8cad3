from rgcn_b200.data import Data  # noqa: F401
