"""Host-side helpers around the path: logposterior, theta construction from params_main.yaml, parameter loading."""
