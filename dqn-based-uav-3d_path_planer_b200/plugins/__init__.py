"""Reference-facing plug-ins.  The reference resolves Env / Agent / Trainer classes by name from XML
(FactoryClass/*.py: importlib.import_module(<Type>) then getattr(module, <Type>)), so every module
here is named after the class it holds.  Put this directory on sys.path and name
the types in the XML files: <Env_Type>PathPlan_City_B200</Env_Type>,
<Trainer_Type>DQN_Trainer_B200 | DDQN_Trainer_B200 | DuelingDQN_Trainer_B200</Trainer_Type>."""
