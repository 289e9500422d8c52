"""lds_gnn_b200 — B200-native (sm_100a) drop-in for the LDS outer-step hot path of andreas-grafberger/lds-gnn.

Layout (mirrors the reference's `src/` tree for the modules on the path):
  csrc/       hand-written CUDA kernels + the C ABI (include/lds_b200.h) -> lib/liblds_b200.so
  kernels.py  ctypes wrappers on torch CUDA tensors
  models/     BernoulliGraphModel, Sampler/sample_graph, MetaDenseGCN, MetaDenseGraphConvolution, factory
  trainers/   OuterProblemTrainer (fused train_step), InnerProblemTrainer.model_forward, BilevelProblemRunner
  utils/      graph tensor utilities, evaluation, early stopping
  config/     sacred-style ingredients + loaders for the reference's sacred JSON / seml YAML LDS configs
"""
__version__ = "0.1.0"
