"""Stand-in for the un-vendored dependency lib5c==0.6.0 (reference
requirements.txt:15).  Only the four functions on the run_to_qvalues path are
provided; they forward to the restatements in oracle/.  Used solely by
oracle/refrun.py to import the UNMODIFIED reference in the build container."""
