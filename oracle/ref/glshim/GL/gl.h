/* Empty stand-in: the reference's main.h pulls in GLEW/GLFW/cuda_gl_interop.h, which want <GL/gl.h>.
 * Nothing on the pathtrace()/denoise() hot path uses OpenGL. (oracle build infrastructure, not product code) */
