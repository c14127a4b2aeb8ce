"""Drop-in `simple_knn` (reference: dgmesh/submodules/simple-knn): `from simple_knn._C import distCUDA2`."""
