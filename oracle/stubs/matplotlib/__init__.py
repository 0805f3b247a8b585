"""Empty stand-in: src/agents/base.py:7 imports matplotlib.pyplot, used only inside Base.test."""
