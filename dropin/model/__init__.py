"""Put `<repo>/dropin` ahead of the reference checkout on PYTHONPATH and the reference's
`from model.egtr import ...` / `from model.deformable_detr import ...` statements
(`/root/reference/evaluate_egtr.py:21-23`) resolve to the B200 implementation."""
