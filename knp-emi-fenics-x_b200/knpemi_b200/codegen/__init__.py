"""RHS -> CUDA code generator for membrane model modules."""
from .build import (BuildError, CODEGEN_VERSION, EmitOptions, RUNTIME_LIB, build_runtime,
                    compile_model, generate, generate_from_source, library_path, model_library)
from .ir import ModelSourceError
from .parse import parse_model_source

__all__ = ["BuildError", "CODEGEN_VERSION", "EmitOptions", "ModelSourceError", "RUNTIME_LIB",
           "build_runtime", "compile_model", "generate", "generate_from_source", "library_path",
           "model_library", "parse_model_source"]
