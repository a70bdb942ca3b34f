"""Drop-in `det3d` namespace: the module paths hydra resolves from the reference's configs
(`_target_: det3d.models....`) re-exported from pillarnext_b200.modules."""
