"""Empty stand-in for cvxpy: the MPC controller is out of scope."""
