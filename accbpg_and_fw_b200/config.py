"""Run-time switches of the package (plain module attributes; set them before calling a driver).

linear_images
    D-opt's Gram matrix M(x) = H diag(x) H^T and the linear-inverse objectives' A x are *linear* in x.  The
    accelerated drivers only ever evaluate f at x0, at fresh prox points z+ and at convex combinations
    (1-theta) a + theta b of points whose image they already hold, so with this switch on they carry the images
    along and form the image of a combination with one axpby instead of another pass over H / A
    ("linearity shortcuts are allowed as optimisations").  Values agree with the
    direct evaluation to rounding (the tests run both ways); the image of x is re-formed from scratch every
    `reanchor_every` iterations so rounding cannot accumulate.  Set to False to evaluate every point from scratch.
"""
linear_images = True
reanchor_every = 64
