"""Agent-parameter halves of the reference presets for the five BASELINE configurations (rl_coach/presets/*.py).

A Coach preset = agent parameters + environment + schedule + graph manager; only the first is on the
replay -> learn path, so each module here defines ``agent_params`` (with this package's Parameters classes, whose
``path`` strings name the device classes) and the observation / action geometry the synthetic benchmarks use.
``INTEGRATION.md`` shows the two-line edit that turns the reference preset into these.
"""
