"""CPU oracle for the DiffMusic guidance hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is product code.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it, and there only as the checker
(or as the timed CPU baseline), never on the shipped path.  The product (``diffmusic_b200``) never imports
this package and fails loudly when its CUDA library is missing.

What is restated here (each function cites the reference file:line it follows, paths under /root/reference):

* ``oracle.ddim_base``  -- diffusers==0.31.0 ``DDIMScheduler`` (requirements.txt:6).  diffusers is NOT installed
  in this image and is not vendored by the reference, so this file restates the published 0.31.0 algorithm.
  **Parity unpinned** for this one third-party base class: the reference holds no test or golden vector for it
  (SURVEY.md section 8c).  Known answers frozen from the restatement: timesteps(500) = 999,997,...,1;
  final_alpha_cumprod = 0.99849999; alphas_cumprod[999] = 1.4230386e-4.
* ``oracle.operators``  -- diffmusic/inverse_problem/{operator,noise}.py on CPU torch/torchaudio (the same library
  calls the reference makes).  **Pinned**: checked against fixtures produced by importing the reference's own
  modules in the build container (tests/golden/make_golden.py -> tests/golden/*.npz).
* ``oracle.steps``      -- diffmusic/schedulers/scheduling_{ddim,dps,mpgd,dsg,diffmusic}.py ``.step``.  **Pinned**
  against fixtures produced by running the reference's unmodified scheduler files (through a diffusers shim whose
  base class is ``oracle.ddim_base``; so the guidance algebra is pinned, the diffusers base is not).
* ``oracle.fad``        -- fadtk/fad.py:41-47, 50-119, 303-350 and fadtk/utils.py:13-46 in NumPy / SciPy float64.
  fadtk cannot be imported (hypy_utils, embedding-model loaders) and its only test needs network models and a
  missing blob (fadtk/stats/fma_pop.npz).  **Pinned**: tests/golden/make_fad_golden.py cuts the reference's own
  functions out of fadtk/fad.py and fadtk/utils.py with ``ast`` and runs them unmodified on seeded fp16 embeddings
  (statistics incl. the fp16 mean of SURVEY.md D.11, Frechet distance, FAD-inf, the online merge over .npy files)
  -> tests/golden/fad.npz; tests/test_oracle_vs_golden.py holds the restatement to those outputs.
"""
