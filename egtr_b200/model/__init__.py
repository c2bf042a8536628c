"""Drop-in mirrors of the reference's `model.egtr` / `model.deformable_detr` modules."""
