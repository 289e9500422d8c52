"""Stand-in for torchmeta==1.2.1 (environment.yml:30). TEST INFRASTRUCTURE ONLY."""
