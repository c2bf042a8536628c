"""Drop-in for the two helpers of `lib/pytorch_misc.py` the evaluation loop imports (train_egtr.py:30): same names, served by
egtr_b200.evaluation."""
from egtr_b200.evaluation import descending_indices as argsort_desc  # noqa: F401
from egtr_b200.evaluation import rows_equal as intersect_2d  # noqa: F401
