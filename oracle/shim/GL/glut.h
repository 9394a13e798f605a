/* No-op GL/GLUT stubs: the reference calls these from its render loop for progress display. */
#ifndef MIRO_ORACLE_GLUT_STUB
#define MIRO_ORACLE_GLUT_STUB
typedef unsigned int GLenum; typedef int GLsizei; typedef float GLfloat; typedef double GLdouble; typedef void GLvoid; typedef unsigned int GLbitfield;
#define GL_COLOR_BUFFER_BIT 0x4000
#define GL_DEPTH_BUFFER_BIT 0x100
#define GL_TRIANGLES 4
#define GL_RGB 0x1907
#define GL_UNSIGNED_BYTE 0x1401
#define GL_BACK 0x405
#define GL_FRONT 0x404
#define GL_PROJECTION 0x1701
#define GL_MODELVIEW 0x1700
#define GL_LINES 1
#define GL_LINE_LOOP 2
static inline void glClear(GLbitfield) {}
static inline void glBegin(GLenum) {}
static inline void glEnd() {}
static inline void glVertex3f(float, float, float) {}
static inline void glColor3f(float, float, float) {}
static inline void glRasterPos2f(float, float) {}
static inline void glDrawPixels(GLsizei, GLsizei, GLenum, GLenum, const GLvoid*) {}
static inline void glFinish() {}
static inline void glFlush() {}
static inline void glutSwapBuffers() {}
static inline void glDrawBuffer(GLenum) {}
static inline void glMatrixMode(GLenum) {}
static inline void glLoadIdentity() {}
static inline void gluPerspective(double, double, double, double) {}
static inline void gluLookAt(double, double, double, double, double, double, double, double, double) {}
static inline void glClearColor(float, float, float, float) {}
static inline void glPolygonMode(GLenum, GLenum) {}
static inline void glPushMatrix() {}
static inline void glPopMatrix() {}
static inline void glTranslatef(float, float, float) {}
#endif
