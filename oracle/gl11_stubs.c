/* No-op definitions of the 18 OpenGL-1.1 entry points the unmodified reference host library links
 * against (they are never reached on the headless render path: SetCudaDevice(0), CUDA atlases,
 * AddRenderBuf).  Test infrastructure for oracle/_ref only. */
#define STUB(name) void name(void) {}
STUB(glBindTexture) STUB(glClear) STUB(glClearColor) STUB(glColorMask) STUB(glDeleteTextures)
STUB(glDepthMask) STUB(glDisable) STUB(glDrawElements) STUB(glEnable) STUB(glFinish)
STUB(glGenTextures) STUB(glGetTexImage) STUB(glGetTexLevelParameteriv)
STUB(glPixelStorei) STUB(glReadPixels) STUB(glTexParameteri) STUB(glViewport)
unsigned int glGetError(void) { return 0; }
