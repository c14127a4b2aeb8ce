"""Drop-in for the part of the third-party `nvdiffrast` package that DG-Mesh uses:
`import nvdiffrast.torch as dr` -> dr.RasterizeGLContext, dr.rasterize, dr.interpolate, dr.antialias, backed
by this package's CUDA triangle rasteriser (meshrast.py / csrc/meshrast.cu).  No OpenGL context is created
(the reference's dr.RasterizeGLContext() at dgmesh/train.py:71 is what fails on a headless node)."""
__version__ = "0.0-dgmesh_b200"
