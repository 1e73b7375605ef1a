"""Mirror of the reference's `model` package (nerf_rpn/model/): same module paths, class names, constructor
arguments, forward() signatures and state_dict keys for the hot path, backed by libnerf_rpn_b200 kernels.
run_rpn.py resolves `from model.xxx import ...` to this package when dropin/ precedes the reference on sys.path
(see INTEGRATION.md)."""
