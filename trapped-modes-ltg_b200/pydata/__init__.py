"""Mirror of the part of the reference's ``pydata`` package that sits directly around the FCD
path: ``analyze.mask`` / ``analyze.center`` (and ``load_image``), running on the GPU."""
