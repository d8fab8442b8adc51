"""Run-time switches of the package (plain module attributes; set them before calling a driver).

linear_images
    D-opt's Gram matrix M(x) = H diag(x) H^T and the linear-inverse objectives' A x are *linear* in x.  The
    accelerated drivers only ever evaluate f at x0, at fresh prox points z+ and at convex combinations
    (1-theta) a + theta b of points whose image they already hold, so with this switch on they carry the images
    along and form the image of a combination with one axpby instead of another pass over H / A
    ("linearity shortcuts are allowed as optimisations").  Values agree with the
    direct evaluation to rounding (the tests run both ways); the image of x is re-formed from scratch every
    `reanchor_every` iterations so rounding cannot accumulate.  Set to False to evaluate every point from scratch.

pipeline
    ABPG, ABDA and BPG without line search take no data-dependent decision inside an iteration except the stopping
    test, so the drivers enqueue iteration k+1 before the scalars of iteration k have been read back (deferred read,
    accbpg_ctx_read_async): the GPU never idles between iterations.  Recorded values and the returned iterate are
    those of the synchronous loop; when the stopping test fires, the one iteration enqueued beyond it is discarded.
    Switched off automatically with verbose=True or restart=True.

peer_allreduce
    Column-sharded runs exchange through NVLink peer memory inside this library's kernels instead of NCCL calls between
    them (buffers from torch.distributed._symmetric_memory): the ranks' Gram matrices are summed behind the SYRK
    (accbpg_dopt_gram_allreduce), and D_opt_FW(_away) exchanges its selection records and the chosen column
    (accbpg_fw_run_peer).  Falls back to NCCL when symmetric memory cannot be set up.

burg_exchange
    Column-sharded Burg-simplex prox as ONE kernel per rank that exchanges the two partial sums of every bisection /
    Newton step with all ranks (accbpg_burg_simplex_prox_peer) instead of gathering gg once and replaying the recurrence
    on the gathered vector on every rank.  Off by default: an NVLink exchange per Newton step costs more than the
    replicated arithmetic at the shapes measured (85 / 130 us against 74 / 94 us per prox on 2 / 8 B200).

fused_small
    BPG on a D-optimal design objective with the Burg kernel on the simplex whose H fits in the shared memory of one SM
    (configs[0], 80 x 200, does): the whole solve is ONE kernel launch (accbpg_dopt_bpg_small) instead of a dozen launches
    and a host decision per line-search trip.  Same control flow and histories; T is stamped evenly over the solve's
    wall time (the host does not see individual iterations).
"""
linear_images = True
reanchor_every = 64
pipeline = True
peer_allreduce = __import__('os').environ.get('ACCBPG_PEER_ALLREDUCE', '1') != '0'
burg_exchange = __import__('os').environ.get('ACCBPG_BURG_EXCHANGE', '0') == '1'
fused_small = __import__('os').environ.get('ACCBPG_FUSED_SMALL', '1') != '0'
