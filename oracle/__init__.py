"""TEST INFRASTRUCTURE (not product code).  CPU oracle for the LaDCast ensemble-rollout hot path:

* `oracle/shim/`           restatement of the diffusers==0.32.1 / xarray names the reference imports, so the
                           UNMODIFIED reference under /root/reference runs on CPU (builder container only);
* `oracle/ladcast_oracle.py` self-contained functional restatement of the path (travels to the GPU box);
* `oracle/make_golden.py`  runs the unmodified reference here and writes tests/golden/*.npz.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this package.
"""
