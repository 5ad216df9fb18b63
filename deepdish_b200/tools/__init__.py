"""Mirror of the hot-path pieces of the reference's ``tools`` package."""
